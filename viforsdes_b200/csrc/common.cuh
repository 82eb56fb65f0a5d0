// Shared declarations for libvisde (sm_100a).  Internal; the public surface is include/visde.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/visde.h"

namespace visde {

void set_error(const char* fmt, ...);

#define VISDE_CUDA_CHECK(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::visde::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                         cudaGetErrorString(_e));                                         \
      return VISDE_ECUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define VISDE_REQUIRE(cond, ...)                                                          \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::visde::set_error(__VA_ARGS__);                                                    \
      return VISDE_EINVAL;                                                                \
    }                                                                                     \
  } while (0)

// "Done once per device" flag for per-kernel attributes (cudaFuncSetAttribute is per device: a process that drives
// several GPUs must opt every one of them into > 48 KB of dynamic shared memory).
struct DeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool needed(int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    return dev >= 64 || !((mask.load(std::memory_order_acquire) >> dev) & 1ull);
  }
  void done(int dev) {
    if (dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
  }
};

constexpr int kGateR = 0, kGateU = 1, kGateN = 2;  // nn.GRU row-block order r, z(update), n
// stash slots per (b, t, layer): r, u, n, n_hh (= W_hn h + b_hn), h_new
constexpr int kStashR = 0, kStashU = 1, kStashN = 2, kStashNhh = 3, kStashH = 4, kStashSlots = 5;
// backward scratch slots per (b, t, layer): d r_pre, d u_pre, d n_pre, d n_hh (= d n_pre * r)
constexpr int kDgSlots = 4;

// Everything the recurrence kernels need.  All pointers are device pointers.
struct PathParams {
  int64_t B, T;
  int S, C, P, H, NL;
  int n_tril, n_out;
  float dt, sqrt_dt;
  const float* x0;      // [B,S]
  const float* theta;   // [B,P]
  const float* eps;     // [B,T,S]
  const float* gi_ctx;  // [B*T,3H]  context rows of W_ih_l0 applied to ctx, + b_ih_l0 (K0 output)
  const float* w_ih[VISDE_MAX_LAYERS];
  const float* w_hh[VISDE_MAX_LAYERS];
  const float* b_ih[VISDE_MAX_LAYERS];
  const float* b_hh[VISDE_MAX_LAYERS];
  const float* out_w;
  const float* out_b;
  // forward outputs
  float* paths;  // [B,T+1,S]
  float* means;  // [B,T,S]
  float* chol;   // [B,T,S,S]
  float* stash;  // [B,T,NL,5,H] or nullptr
  float* raw;    // [B,T,n_tril] raw (unfloored) Cholesky params, or nullptr
  // backward inputs
  const float* g_paths;  // [B,T+1,S]
  const float* g_means;  // [B,T,S]
  const float* g_chol;   // [B,T,S,S]
  // backward outputs
  float* grad_x0;  // [B,S]
  float* dg;       // [B*T,NL,4,H]
  float* dout;     // [B*T,n_out]
  float* sdg;      // [B,3H]  sum_t d_gi of layer 0
  float* cta_part; // fast family: per-CTA partial sums of the thin weight-gradient pieces (or nullptr)
  // wide-state tensor-core family (path_tcw.cu): weight tile images, tiled noise, tiled step records, tiled cotangents
  void* wimg;
  float* epst;   // [tile][t][S][128]
  float* otile;  // [tile][t][2S + S(S+1)/2][128]: mu | raw Cholesky | z_{t+1}   (lives in the stash: the backward reads it)
  float* ctile;  // [tile][t][3S + S*S][128]: gP[t+1] | gM | gL | eps
};

__host__ __device__ inline int64_t stash_row_floats(int NL, int H) { return (int64_t)NL * kStashSlots * H; }

// MUFU-based gates on the serial critical path: raw ex2.approx / rcp.approx (~1 ulp each, no range
// guards, no IEEE-division slow path); absolute error ~2e-7, which is what the hidden state needs.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_f(float x) {
  return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f);
}
// packed dual-fp32 FMA (sm_100 FFMA2): d.x += a.x * b.x ; d.y += a.y * b.y
__device__ __forceinline__ void fma2(float2& d, const float2& a, const float2& b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}

// opt-in stage profiler (api.cu)
struct StageTimer {
  int rec;
  StageTimer(int stage, int launches, cudaStream_t st);
  ~StageTimer();
  cudaStream_t st_;
};

// launchers (return 0 / VISDE_E*)
int launch_path_fwd_generic(const PathParams& p, cudaStream_t st);
int launch_path_bwd_generic(const PathParams& p, cudaStream_t st);
bool fast_supported(const PathParams& p);
int launch_path_fwd_fast(const PathParams& p, cudaStream_t st);
int launch_path_bwd_fast(const PathParams& p, cudaStream_t st);
// batch-tiled family for B > #SM (path_tiled.cu): NB trajectories per CTA, weights in shared memory
int tiled_batch_tile(int64_t B, bool force, bool bwd);
int launch_path_fwd_tiled(const PathParams& p, int NB, cudaStream_t st);
int launch_path_bwd_tiled(const PathParams& p, int NB, int* ncta, cudaStream_t st);  // same outputs + partial records as the fast family
// register-resident family for wide state spaces, 4 < S <= 16 (path_fast_s.cu); the backward also writes the bias
// gradients (p.cta_part: fasts_partials_floats()); dW_ih_l0[:, :S], dW_out, db_out are left to the GEMM stage
bool fasts_supported(const PathParams& p);
size_t fasts_partials_floats(int NL, int H);
int launch_path_fwd_fasts(const PathParams& p, cudaStream_t st);
int launch_path_bwd_fasts(const PathParams& p, const visde_weight_grads* gw, cudaStream_t st);
// dW_ih_l0[:, :S], dW_out, db_out of the wide-state family in one time-parallel pass over d_pre / d_out / the stash (fasts_thin.cu)
size_t fasts_thin_partial_floats(int64_t B, int64_t T, int S, int H);
int launch_fasts_thin_grads(const PathParams& p, const visde_weight_grads* gw, float* partials, size_t partial_floats,
                            cudaStream_t st);
// tensor-core recurrence family for large batches (path_tc.cu): 128 trajectories per CTA, tcgen05 gate GEMMs
bool tc_rec_supported(const PathParams& p);
int launch_gth(const PathParams& p, float* gth, cudaStream_t st);  // per-trajectory constant part of the layer-0 gates
int launch_path_fwd_tc(const PathParams& p, cudaStream_t st);      // expects tiled gi_ctx that already includes gth
// BPTT on tensor cores: p.stash / p.dg / p.dout ([tile][t][16][128]) are tiled buffers, p.grad_x0 per-trajectory
int launch_path_bwd_tc(const PathParams& p, cudaStream_t st);
// biases, dW_ih_l0[:, :S], dW_out, db_out and p.sdg from the tiled dg / dout / stash (per-tile partials + fixed-order sum)
size_t tc_thin_partial_floats(int64_t B, int NL, int S);
int launch_tc_thin_grads(const PathParams& p, const float* dout_tiled, const visde_weight_grads* gw, float* partials,
                         size_t partial_floats, cudaStream_t st);
// wide-state tensor-core family (4 < S <= 10, NL = 2): path_tcw.cu forward, path_tcw_bwd.cu backward
constexpr int kTcwMaxS = 10;  // S (S + 1) / 2 <= 64: the Cholesky rows of W_out are one N = 64 MMA
// row-fastest tiled per-step record written by the wide forward: mu (S) | raw Cholesky entries (S (S + 1) / 2) | z_{t+1} (S)
__host__ __device__ constexpr int tcw_out_feats(int S) { return 2 * S + S * (S + 1) / 2; }
// row-fastest tiled cotangent record read by the wide backward: gP[t+1] (S) | gM (S) | gL (S x S) | eps (S)
__host__ __device__ constexpr int tcw_cot_feats(int S) { return 3 * S + S * S; }
bool tcw_rec_supported(const PathParams& p);
size_t tcw_image_bytes();
int launch_tcw_images(const PathParams& p, void* img, bool fwd, bool bwd, cudaStream_t st);
int launch_tcw_tile(const float* src, int64_t B, int64_t T, int F, int64_t bstride, int64_t tstride, float* dst, int FD, int f_off,
                    cudaStream_t st);
int launch_tcw_tile_multi(const float* const* src, const int64_t* bstride, const int64_t* tstride, const int* F, const int* f_off,
                          int nsrc, int64_t B, int64_t T, float* dst, int FD, cudaStream_t st);
int launch_path_fwd_tcw(const PathParams& p, cudaStream_t st);  // needs p.wimg (forward images), p.epst, p.otile, tiled gi_ctx
int launch_path_bwd_tcw(const PathParams& p, cudaStream_t st);  // needs p.wimg (backward images), p.ctile, p.otile, tiled stash
size_t tcw_thin_partial_floats(int64_t B, int NL, int S);
int launch_tcw_thin_grads(const PathParams& p, const visde_weight_grads* gw, float* partials, size_t partial_floats,
                          cudaStream_t st);
// [ceil(B/128)][T][F][128] row-fastest -> [B][T][F]
int launch_untile(const float* in, float* out, int64_t B, int64_t T, int F, cudaStream_t st);
size_t fast_partials_floats(int NL, int H, int S);
// biases, dW_ih_l0[:, :S], dW_out, db_out from the per-CTA partials written by path_bwd_fast
int launch_fast_partials_reduce(const PathParams& p, const visde_weight_grads* gw, cudaStream_t st, int ncta = 0);

// --- SIMT fp32 GEMMs with (b,t) row gathering -------------------------------------------
// A "row source": row k = (b, t) with b = k / T, t = k % T lives at
//   base + b*bstride + (t + tshift)*tstride + col; rows with t + tshift < 0 read as zero.
// base == nullptr means a column of ones (used to fold bias gradients into a GEMM).
struct RowSrc {
  const void* base;
  int64_t bstride, tstride;
  int tshift;
  int ncols;
  int dtype;  // VISDE_F32 / VISDE_BF16
};

// out[k, n] = bias[n] + sum_c A[k, c] * W[n*ldw + c]          (k over B*T rows, "NT")
int launch_gemm_nt(const RowSrc& A, int64_t B, int64_t T, int K, const float* W, int ldw, int N,
                   const float* bias, float* out, int ldo, cudaStream_t st);
// out[(b,t), c] = sum_n A[k, n] * W[n*ldw + c]   written to a strided fp32/bf16 view ("NN")
int launch_gemm_nn(const RowSrc& A, int64_t B, int64_t T, int K, const float* W, int ldw, int N,
                   void* out, int64_t out_bstride, int64_t out_tstride, int out_dtype,
                   cudaStream_t st);
// C[m, n] = sum_k A[k, amap(m)] * Bcat[k, n]; Bcat = concatenation of up to 3 row sources.
// amap: m < a_split -> m, else m + a_skip (selects the (r,u,n_hh) columns of a 4H-wide row).
// Split-K over CTAs into `partials`, then a fixed-order reduction writes C (ldc) -- deterministic.
struct TnOut {
  float* ptr;  // destination for columns [col0, col0+ncols) of the product; row-major, ld
  int ld;
  int col0, ncols;
};
int launch_gemm_tn(const RowSrc& A, int M, int a_split, int a_skip, const RowSrc* Bsrc, int nsrc,
                   int64_t B, int64_t T, const TnOut* outs, int nouts, float* partials,
                   size_t partial_floats, cudaStream_t st);
size_t gemm_tn_partial_floats(int M, int N, int64_t K);
// grad_theta = sdg . W_ih_l0[:, col0:col0+P] and (dw != nullptr) dW_ih_l0[:, col0:col0+P] = sdg^T theta in one small launch
bool theta_grads_supported(int P);
int launch_theta_grads(const float* sdg, const float* theta, const float* w_ih0, int64_t B, int G, int P, int ld0, int col0,
                       float* grad_theta, float* dw, cudaStream_t st);

// --- tensor-core (tcgen05, 3xTF32) versions of K0 / K3 / K4 (tc_gemm.cu) ------------------------
bool tc_supported(int H, int NL, int C, const visde_ctx_view* ctx);
size_t tc_weight_scratch_floats(int H, int C);
size_t tc_wgrad_partial_floats(int NL, int C);
int tc_split_weights(const float* w_ih0, int ld0, int S, int H, int C, float* scratch, cudaStream_t st);
// bias: per-column [3H] or nullptr.  tiled: row tiles are 128 trajectories at one grid step and gi_ctx is written
// row-fastest, [ceil(B/128)][T][3H][128] (+ rowbias_tiled [ceil(B/128)][3H][128] per trajectory), for path_tc.cu
int tc_ctx_proj(const visde_ctx_view* ctx, int64_t B, int64_t T, int C, int H, const float* wsplit, const float* bias,
                const float* rowbias_tiled, float* gi_ctx, bool tiled, cudaStream_t st);
int tc_grad_ctx(const float* dg, int64_t dg_row, int64_t B, int64_t T, int C, int H, const float* wsplit,
                const visde_ctx_grad_view* out, bool dg_tiled, cudaStream_t st);
int tc_wgrads(const visde_ctx_view* ctx, const float* dg, const float* stash, int64_t B, int64_t T, int S, int C, int P,
              int H, int NL, const visde_weight_grads* gw, float* partials, size_t partial_floats, cudaStream_t st);

// same products from the row-fastest tiled dg / stash of the tensor-core family
int tc_wgrads_tiled(const visde_ctx_view* ctx, const float* dg, const float* stash, int64_t B, int64_t T, int S, int C,
                    int P, int H, int NL, const visde_weight_grads* gw, float* partials, size_t partial_floats,
                    cudaStream_t st);

// --- ELBO ---------------------------------------------------------------------------------
struct ElboParams {
  int64_t B, T;
  int S, P;
  int sde_kind;
  uint32_t pos_mask;
  float dt;
  const float* z;
  const float* means;
  const float* chol;
  const float* theta;
  const float* drift;
  const float* diffusion;
  visde_obs obs;
  float* terms;         // fwd out [B,4]
  const float* g_terms; // bwd in  [B,4]
  float* g_z;
  float* g_means;
  float* g_chol;
  float* g_theta;
  float* g_drift;
  float* g_diffusion;
  int vec16;  // set by the launcher: the S x S factor blocks may be accessed with 16-byte loads / stores
};
int launch_elbo_fwd(const ElboParams& p, cudaStream_t st);
int launch_elbo_bwd(const ElboParams& p, cudaStream_t st);

}  // namespace visde
