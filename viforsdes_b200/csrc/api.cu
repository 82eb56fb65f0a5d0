// extern "C" surface of libvisde (include/visde.h): validation, workspace carving, kernel
// sequencing.  No torch types; every buffer is caller-owned except inside visde_session.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace visde {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- opt-in profiler ------------------------------------------------------------------------
namespace {
struct ProfRecord {
  cudaEvent_t beg, end;
  int stage, launches;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRecord> g_prof;
size_t g_prof_used = 0;
}  // namespace

StageTimer::StageTimer(int stage, int launches, cudaStream_t st) : rec(-1), st_(st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on || g_prof_used >= g_prof.size()) return;
  rec = (int)g_prof_used++;
  g_prof[rec].stage = stage;
  g_prof[rec].launches = launches;
  cudaEventRecord(g_prof[rec].beg, st);
}
StageTimer::~StageTimer() {
  if (rec >= 0) cudaEventRecord(g_prof[rec].end, st_);
}

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// fp32 copy of a bf16 context (reserved whenever the dims are eligible for the tcgen05 GEMM stages)
size_t ctx_f32_bytes(const visde_dims* d) {
  const bool dims_ok = d->H == 64 && d->NL <= 2 && (d->C == 128 || d->C == 256);
  return dims_ok ? align_up(sizeof(float) * (size_t)d->B * d->T * d->C) : 0;
}

int check_dims(const visde_dims* d) {
  VISDE_REQUIRE(d != nullptr, "dims is NULL");
  VISDE_REQUIRE(d->B >= 0 && d->T >= 0, "B and T must be non-negative (got %lld, %lld)", (long long)d->B,
                (long long)d->T);
  VISDE_REQUIRE(d->S >= 1 && d->S <= VISDE_MAX_STATE, "state_dim must be in [1, %d], got %d", VISDE_MAX_STATE, d->S);
  VISDE_REQUIRE(d->NL >= 1 && d->NL <= VISDE_MAX_LAYERS, "num_layers must be in [1, %d], got %d", VISDE_MAX_LAYERS,
                d->NL);  // models/head.py:33-36
  VISDE_REQUIRE(d->H >= 1 && d->H <= VISDE_MAX_HIDDEN, "hidden_dim must be in [1, %d], got %d", VISDE_MAX_HIDDEN, d->H);
  VISDE_REQUIRE(d->C >= 0 && d->P >= 0, "context_dim and sde_param_dim must be non-negative");
  return VISDE_OK;
}

struct StashLayout {
  size_t h_floats, raw_floats, otile_floats;
};
// B rounded up to the 128-trajectory tile of the tensor-core recurrence (its stash / gi_ctx layouts are tiled)
size_t padded_B(const visde_dims* d) { return ((size_t)d->B + 127) / 128 * 128; }

bool tcw_rec_possible(const visde_dims* d);
StashLayout stash_layout(const visde_dims* d) {
  // the wide-state tensor-core family keeps its tiled step records (mu | raw Cholesky | z) for the backward
  const size_t ot = tcw_rec_possible(d) ? padded_B(d) * d->T * (size_t)tcw_out_feats(d->S) : 0;
  return {padded_B(d) * d->T * d->NL * kStashSlots * d->H, (size_t)d->B * d->T * (size_t)(d->S * (d->S + 1) / 2), ot};
}

// AUTO switches to the tensor-core recurrence family from this batch size on: below it the 128-trajectory tiles
// leave most SMs idle and the fp32 SIMT families (one / eight trajectories per CTA) finish sooner (measured
// break-even for fwd + bwd on 148 SMs is B ~ 2 600)
constexpr int64_t kTcRecMinBatch = 3072;

// dims-only version of use_tc_rec (workspace sizing; the run-time decision also looks at the context view)
bool tc_rec_possible(const visde_dims* d) {
  const int fam = d->variant & 0xff;
  return (fam == VISDE_VARIANT_TC || (fam == VISDE_VARIANT_AUTO && d->B >= kTcRecMinBatch)) && d->H == 64 && d->NL <= 2 &&
         d->S <= 4 && (d->C == 128 || d->C == 256) && !(d->variant & VISDE_FLAG_NO_TENSOR_CORES);
}

// wide-state tensor-core recurrence (path_tcw.cu): 4 < S <= 10, two layers.  Its alternative is the one-trajectory-per-CTA FAST-S
// family (no batch-tiled family for wide states), so it pays off earlier than the narrow one: measured at S = 10, T = 100
// (profiles/r2_config5.md, 64-row tiles with four threads per row) 768 trajectories 2.56 ms vs 2.38 ms FAST-S, 1 024: 2.68 vs 2.80,
// 1 536: 2.99 vs 4.26, 2 048: 3.27 vs 5.40
constexpr int64_t kTcwRecMinBatch = 1024;
bool tcw_rec_possible(const visde_dims* d) {
  const int fam = d->variant & 0xff;
  return (fam == VISDE_VARIANT_TC || (fam == VISDE_VARIANT_AUTO && d->B >= kTcwRecMinBatch)) && d->H == 64 && d->NL == 2 &&
         d->S > 4 && d->S <= kTcwMaxS && (d->C == 128 || d->C == 256) && !(d->variant & VISDE_FLAG_NO_TENSOR_CORES);
}

struct BwdWs {
  size_t dg, dout, sdg, partials, wsplit, cta_part, dg_tiled, ctx_f32, wimg, ctile, total, partial_floats;
};
BwdWs bwd_ws(const visde_dims* d) {
  BwdWs w{};
  size_t bt = (size_t)d->B * d->T;
  int n_out = d->S + d->S * (d->S + 1) / 2;
  int G = 3 * d->H;
  size_t off = 0;
  w.dg = off;
  off += align_up(sizeof(float) * bt * d->NL * kDgSlots * d->H);
  w.dout = off;
  {
    size_t fl = bt * n_out;
    if (tc_rec_possible(d) && padded_B(d) * d->T * 16 > fl) fl = padded_B(d) * d->T * 16;  // tiled [tile][t][16][128]
    if (tcw_rec_possible(d)) fl = padded_B(d) * d->T * n_out;                              // tiled [tile][t][n_out][128]
    off += align_up(sizeof(float) * fl);
  }
  w.sdg = off;
  off += align_up(sizeof(float) * (size_t)d->B * G);
  int64_t K = d->B * d->T;
  size_t pf = gemm_tn_partial_floats(G, d->S + d->C + d->P + 1, K);
  size_t pf2 = gemm_tn_partial_floats(G, d->H + 1, K);
  size_t pf3 = gemm_tn_partial_floats(n_out, d->H + 1, K);
  if (pf2 > pf) pf = pf2;
  if (pf3 > pf) pf = pf3;
  if (d->H == 64 && d->NL <= 2 && (d->C == 128 || d->C == 256)) {
    size_t pt = tc_wgrad_partial_floats(d->NL, d->C);
    if (pt > pf) pf = pt;
  }
  if (tc_rec_possible(d)) {
    size_t pt = tc_thin_partial_floats(d->B, d->NL, d->S);
    if (pt > pf) pf = pt;
  }
  if (d->S > 4 && d->H <= 64) {  // wide-state family
    size_t pt = fasts_thin_partial_floats(d->B, d->T, d->S, d->H);
    if (pt > pf) pf = pt;
  }
  if (tcw_rec_possible(d)) {
    size_t pt = tcw_thin_partial_floats(d->B, d->NL, d->S);
    if (pt > pf) pf = pt;
  }
  w.partial_floats = pf;
  w.partials = off;
  off += align_up(sizeof(float) * pf);
  w.wsplit = off;
  off += align_up(sizeof(float) * tc_weight_scratch_floats(d->H, d->C));
  w.cta_part = off;
  off += align_up(sizeof(float) * fast_partials_floats(d->NL, d->H, d->S));
  w.dg_tiled = off;   // d_pre of the tensor-core backward, row-fastest tiled
  if (tc_rec_possible(d) || tcw_rec_possible(d)) off += align_up(sizeof(float) * padded_B(d) * d->T * d->NL * kDgSlots * d->H);
  w.ctx_f32 = off;    // fp32 copy of a bf16 context for the tcgen05 GEMM stages
  off += ctx_f32_bytes(d);
  w.wimg = off;       // wide-state tensor-core family: weight tile images, tiled cotangent records
  if (tcw_rec_possible(d)) off += tcw_image_bytes();
  w.ctile = off;
  if (tcw_rec_possible(d)) off += align_up(sizeof(float) * padded_B(d) * d->T * (size_t)tcw_cot_feats(d->S));
  w.total = off;
  return w;
}

void fill_common(PathParams& p, const visde_dims* d, float dt, const visde_weights* w) {
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.T = d->T;
  p.S = d->S;
  p.C = d->C;
  p.P = d->P;
  p.H = d->H;
  p.NL = d->NL;
  p.n_tril = d->S * (d->S + 1) / 2;
  p.n_out = d->S + p.n_tril;
  p.dt = dt;
  p.sqrt_dt = sqrtf(dt);
  for (int k = 0; k < d->NL; ++k) {
    p.w_ih[k] = w->w_ih[k];
    p.w_hh[k] = w->w_hh[k];
    p.b_ih[k] = w->b_ih[k];
    p.b_hh[k] = w->b_hh[k];
  }
  p.out_w = w->out_w;
  p.out_b = w->out_b;
}

int check_weights(const visde_dims* d, const visde_weights* w) {
  VISDE_REQUIRE(w != nullptr, "weights is NULL");
  for (int k = 0; k < d->NL; ++k)
    VISDE_REQUIRE(w->w_ih[k] && w->w_hh[k] && w->b_ih[k] && w->b_hh[k], "weights of layer %d are NULL", k);
  VISDE_REQUIRE(w->out_w && w->out_b, "out_proj weights are NULL");
  return VISDE_OK;
}

bool use_tc(const visde_dims* d, const visde_ctx_view* ctx) {
  if (d->variant & VISDE_FLAG_NO_TENSOR_CORES) return false;
  return tc_supported(d->H, d->NL, d->C, ctx);
}

// tensor-core recurrence: explicit request, or AUTO once the batch is large enough that 128-trajectory tiles
// beat the fp32 SIMT families (needs the tcgen05 K0, which folds the per-trajectory constants into gi_ctx)
bool use_tc_rec(const visde_dims* d, const PathParams& p, const visde_ctx_view* ctx) {
  return tc_rec_possible(d) && tc_rec_supported(p) && use_tc(d, ctx);
}
bool use_tcw_rec(const visde_dims* d, const PathParams& p, const visde_ctx_view* ctx) {
  return tcw_rec_possible(d) && tcw_rec_supported(p) && use_tc(d, ctx);
}

size_t fwd_gi_bytes(const visde_dims* d) { return align_up(sizeof(float) * padded_B(d) * d->T * 3 * d->H); }
size_t fwd_wsplit_bytes(const visde_dims* d) { return align_up(sizeof(float) * tc_weight_scratch_floats(d->H, d->C)); }
size_t fwd_gth_bytes(const visde_dims* d) { return align_up(sizeof(float) * padded_B(d) * 3 * d->H); }
// weight images + tiled noise + (inference mode: no stash) the tiled step records of the wide-state tensor-core family
size_t fwd_tcw_bytes(const visde_dims* d) {
  if (!tcw_rec_possible(d)) return 0;
  return tcw_image_bytes() + align_up(sizeof(float) * padded_B(d) * d->T * d->S) +
         align_up(sizeof(float) * padded_B(d) * d->T * (size_t)tcw_out_feats(d->S));
}

bool use_fast(const visde_dims* d, const PathParams& p) {
  if ((d->variant & 0xff) == VISDE_VARIANT_GENERIC) return false;
  return fast_supported(p);
}
// wide-state version of the register-resident family (4 < S <= 16)
bool use_fasts(const visde_dims* d, const PathParams& p) {
  const int fam = d->variant & 0xff;
  return (fam == VISDE_VARIANT_AUTO || fam == VISDE_VARIANT_FAST) && fasts_supported(p);
}

// bf16 context (what the reference's autocast encoder emits, inference/training_context.py:104) -> dense fp32
// [B,T,C] in the workspace, so that the tcgen05 GEMM stages (fp32 TMA tiles, 3xTF32) serve it too; bf16 values are
// exact in fp32, so this is the arithmetic of kernels/forward.py (load, widen, fp32 math)
__global__ void ctx_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, int64_t bstride, int64_t tstride, int64_t B,
                                       int64_t T, int C, float* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*T*C/2 pairs
  const int64_t half_c = C / 2;
  if (idx >= B * T * half_c) return;
  const int64_t c2 = idx % half_c, bt = idx / half_c, t = bt % T, b = bt / T;
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + b * bstride + t * tstride + 2 * c2);
  *reinterpret_cast<float2*>(dst + bt * C + 2 * c2) = __bfloat1622float2(v);
}
// the same widening, eight values per thread (one 16-byte load, two 16-byte stores): rows whose start and strides are 16-byte aligned
__global__ void ctx_bf16_to_f32_vec8_kernel(const __nv_bfloat16* __restrict__ src, int64_t bstride, int64_t tstride, int64_t B,
                                            int64_t T, int C, float* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*T*C/8 octets
  const int64_t oct_c = C / 8;
  if (idx >= B * T * oct_c) return;
  const int64_t c8 = idx % oct_c, bt = idx / oct_c, t = bt % T, b = bt / T;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + b * bstride + t * tstride + 8 * c8));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float f[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // a bf16 is the upper half of the fp32 with the same value
    f[2 * q] = __uint_as_float(w[q] << 16);
    f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
  float4* o = reinterpret_cast<float4*>(dst + bt * C + 8 * c8);
  o[0] = make_float4(f[0], f[1], f[2], f[3]);
  o[1] = make_float4(f[4], f[5], f[6], f[7]);
}
bool ctx_convertible(const visde_dims* d, const visde_ctx_view* ctx) {
  if (!ctx || ctx->dtype != VISDE_BF16 || (d->variant & VISDE_FLAG_NO_TENSOR_CORES)) return false;
  if ((reinterpret_cast<uintptr_t>(ctx->ptr) & 3) || (ctx->batch_stride & 1) || (ctx->time_stride & 1)) return false;
  visde_ctx_view probe{reinterpret_cast<const void*>(uintptr_t(256)), d->T * (int64_t)d->C, d->C, VISDE_F32};
  return tc_supported(d->H, d->NL, d->C, &probe);
}
int convert_ctx(const visde_dims* d, const visde_ctx_view* ctx, float* buf, visde_ctx_view* out, cudaStream_t st) {
  if (d->C % 8 == 0 && (reinterpret_cast<uintptr_t>(ctx->ptr) & 15) == 0 && ctx->batch_stride % 8 == 0 && ctx->time_stride % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(buf) & 15) == 0) {
    const int64_t n8 = d->B * d->T * (d->C / 8);
    ctx_bf16_to_f32_vec8_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(ctx->ptr),
                                                                              ctx->batch_stride, ctx->time_stride, d->B, d->T, d->C, buf);
    VISDE_CUDA_CHECK(cudaGetLastError());
    *out = visde_ctx_view{buf, d->T * (int64_t)d->C, d->C, VISDE_F32};
    return VISDE_OK;
  }
  const int64_t n = d->B * d->T * (d->C / 2);
  ctx_bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(ctx->ptr),
                                                                       ctx->batch_stride, ctx->time_stride, d->B, d->T, d->C, buf);
  VISDE_CUDA_CHECK(cudaGetLastError());
  *out = visde_ctx_view{buf, d->T * (int64_t)d->C, d->C, VISDE_F32};
  return VISDE_OK;
}

__global__ void fill_loss_cotangent_kernel(float* g_terms, int64_t B) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // loss = -mean_b(obs + sde - gen + jac)   (evidence_lower_bound.py:64, trainer.py:197)
  const float s = -1.0f / (float)B;
  g_terms[b * 4 + 0] = s;
  g_terms[b * 4 + 1] = s;
  g_terms[b * 4 + 2] = -s;
  g_terms[b * 4 + 3] = s;
}

__global__ void add_inplace_kernel(float* dst, const float* src, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// user-SDE sessions: x[b,t,:] = to_state(z[b,t,:]) for t < T (inference/state_space.py:20-25; softplus threshold 20)
__global__ void state_rows_kernel(const float* __restrict__ z, int64_t B, int64_t T, int S, uint32_t mask, float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T * S) return;
  const int s = (int)(i % S);
  const int64_t bt = i / S, b = bt / T, t = bt % T;
  const float v = z[(b * (T + 1) + t) * S + s];
  x[i] = ((mask >> s) & 1u) ? (v > 20.f ? v : log1pf(expf(v))) : v;
}
// g_z[b,t,:] += g_x[b,t,:] * d to_state / dz for t < T (the chain the reference's autograd walks from drift / diffusion)
__global__ void chain_state_grad_kernel(const float* __restrict__ z, const float* __restrict__ g_x, int64_t B, int64_t T, int S,
                                        uint32_t mask, float* __restrict__ g_z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T * S) return;
  const int s = (int)(i % S);
  const int64_t bt = i / S, b = bt / T, t = bt % T;
  const int64_t zi = (b * (T + 1) + t) * S + s;
  const float v = z[zi];
  const float d = ((mask >> s) & 1u) ? (v > 20.f ? 1.f : 1.f / (1.f + expf(-v))) : 1.f;
  g_z[zi] += g_x[i] * d;
}

}  // namespace
}  // namespace visde

using namespace visde;

extern "C" {

int visde_version(void) { return VISDE_VERSION; }
const char* visde_last_error(void) { return g_err; }

size_t visde_stash_bytes(const visde_dims* d) {
  if (check_dims(d) != VISDE_OK) return 0;
  StashLayout s = stash_layout(d);
  return align_up(sizeof(float) * s.h_floats) + align_up(sizeof(float) * s.raw_floats) + align_up(sizeof(float) * s.otile_floats) + 256;
}

size_t visde_workspace_bytes(const visde_dims* d, int backward) {
  if (check_dims(d) != VISDE_OK) return 0;
  if (!backward) return fwd_gi_bytes(d) + fwd_wsplit_bytes(d) + fwd_gth_bytes(d) + ctx_f32_bytes(d) + fwd_tcw_bytes(d) + 256;
  return bwd_ws(d).total + 256;
}

int visde_recurrence_family(const visde_dims* d, int backward) {
  int rc = check_dims(d);
  if (rc) return rc;
  PathParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.T = d->T; p.S = d->S; p.C = d->C; p.P = d->P; p.H = d->H; p.NL = d->NL;
  p.n_tril = d->S * (d->S + 1) / 2;
  p.n_out = d->S + p.n_tril;
  // the same predicates, in the same order, as visde_path_fwd / visde_path_bwd below
  if (d->T > 0 && tc_rec_possible(d) && tc_rec_supported(p)) return VISDE_FAMILY_TC;
  if (d->T > 0 && tcw_rec_possible(d) && tcw_rec_supported(p)) return VISDE_FAMILY_TC;
  if (use_fasts(d, p)) return VISDE_FAMILY_FAST_S;
  if (use_fast(d, p)) {
    const int fam = d->variant & 0xff;
    const bool pinned = fam == VISDE_VARIANT_FAST || fam == VISDE_VARIANT_TC || (backward && d->T == 0);
    const int nb = pinned ? 0 : tiled_batch_tile(d->B, fam == VISDE_VARIANT_TILED, backward != 0);
    return nb == 8 ? VISDE_FAMILY_TILED8 : nb == 4 ? VISDE_FAMILY_TILED4 : VISDE_FAMILY_FAST;
  }
  return VISDE_FAMILY_GENERIC;
}

int visde_path_fwd(const visde_dims* d, float dt, const float* x0, const visde_ctx_view* ctx,
                   const float* theta, const float* eps, const visde_weights* w, float* paths,
                   float* means, float* chol, void* stash, void* workspace, size_t workspace_bytes,
                   void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_weights(d, w))) return rc;
  VISDE_REQUIRE(dt > 0.f, "time_step must be positive");
  if (d->B == 0) return VISDE_OK;
  VISDE_REQUIRE(x0 && paths, "x0 / paths is NULL");
  VISDE_REQUIRE(d->T == 0 || (ctx && ctx->ptr && eps && means && chol), "NULL tensor argument");
  VISDE_REQUIRE(d->P == 0 || theta, "theta is NULL");
  VISDE_REQUIRE(((d->variant & 0xff) != VISDE_VARIANT_FAST && (d->variant & 0xff) != VISDE_VARIANT_TILED) ||
                    (d->H <= 64 && d->H % 4 == 0 && d->NL <= 2 && d->S <= ((d->variant & 0xff) == VISDE_VARIANT_FAST ? 16 : 4)),
                "fast variant requested for an unsupported shape (H=%d NL=%d S=%d)", d->H, d->NL, d->S);
  VISDE_REQUIRE((d->variant & 0xff) != VISDE_VARIANT_TC ||
                    (d->H == 64 && ((d->NL <= 2 && d->S <= 4) || (d->NL == 2 && d->S <= kTcwMaxS)) &&
                     (use_tc(d, ctx) || ctx_convertible(d, ctx))),
                "tensor-core recurrence requested for an unsupported shape (needs H=64 and NL<=2, S<=4 or NL=2, S<=10; fp32 "
                "16-byte aligned context with C in {128, 256}; got H=%d NL=%d S=%d C=%d)", d->H, d->NL, d->S, d->C);
  if (workspace_bytes < visde_workspace_bytes(d, 0) - 256 || (!workspace && d->T > 0)) {
    set_error("path_fwd: workspace too small (%zu < %zu)", workspace_bytes, visde_workspace_bytes(d, 0));
    return VISDE_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  visde_ctx_view ctx_f32;
  if (d->T > 0 && ctx_convertible(d, ctx)) {
    float* buf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + fwd_gi_bytes(d) + fwd_wsplit_bytes(d) + fwd_gth_bytes(d));
    if ((rc = convert_ctx(d, ctx, buf, &ctx_f32, st))) return rc;
    ctx = &ctx_f32;
  }
  PathParams p;
  fill_common(p, d, dt, w);
  p.x0 = x0;
  p.theta = theta;
  p.eps = eps;
  p.paths = paths;
  p.means = means;
  p.chol = chol;
  float* gi = reinterpret_cast<float*>(workspace);
  p.gi_ctx = gi;
  if (stash) {
    StashLayout s = stash_layout(d);
    p.stash = reinterpret_cast<float*>(stash);
    p.raw = reinterpret_cast<float*>(reinterpret_cast<char*>(stash) + align_up(sizeof(float) * s.h_floats));
    p.otile = reinterpret_cast<float*>(reinterpret_cast<char*>(p.raw) + align_up(sizeof(float) * s.raw_floats));
  }
  const bool tcwrec = d->T > 0 && use_tcw_rec(d, p, ctx);
  if (tcwrec) {
    char* tcw_ws = reinterpret_cast<char*>(workspace) + fwd_gi_bytes(d) + fwd_wsplit_bytes(d) + fwd_gth_bytes(d) + ctx_f32_bytes(d);
    p.wimg = tcw_ws;
    p.epst = reinterpret_cast<float*>(tcw_ws + tcw_image_bytes());
    // inference mode keeps no stash: the step records then live in the workspace
    if (!stash) p.otile = reinterpret_cast<float*>(tcw_ws + tcw_image_bytes() + align_up(sizeof(float) * padded_B(d) * d->T * d->S));
  }
  const bool tcrec = (d->T > 0 && use_tc_rec(d, p, ctx)) || tcwrec;
  if (d->T > 0) {
    // K0: context rows of W_ih_l0 as one time-parallel GEMM, b_ih_l0 folded in
    StageTimer tm(VISDE_STAGE_K0_CTX_GEMM, tcrec ? 3 : 1, st);
    if (use_tc(d, ctx)) {
      float* wsplit = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + fwd_gi_bytes(d));
      rc = tc_split_weights(w->w_ih[0], d->S + d->C + d->P, d->S, d->H, d->C, wsplit, st);
      if (rc) return rc;
      if (tcrec) {
        // the tensor-core recurrence takes every per-trajectory constant of the layer-0 gates (theta columns,
        // b_ih_l0, b_hh_l0[r, u]) from gi_ctx: K0 adds them as a row bias
        float* gth = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + fwd_gi_bytes(d) + fwd_wsplit_bytes(d));
        rc = launch_gth(p, gth, st);
        if (rc) return rc;
        rc = tc_ctx_proj(ctx, d->B, d->T, d->C, d->H, wsplit, nullptr, gth, gi, true, st);
      } else {
        rc = tc_ctx_proj(ctx, d->B, d->T, d->C, d->H, wsplit, w->b_ih[0], nullptr, gi, false, st);
      }
    } else {
      RowSrc A{ctx->ptr, ctx->batch_stride, ctx->time_stride, 0, d->C, ctx->dtype};
      rc = launch_gemm_nt(A, d->B, d->T, d->C, w->w_ih[0] + d->S, d->S + d->C + d->P, 3 * d->H, w->b_ih[0], gi,
                          3 * d->H, st);
    }
    if (rc) return rc;
  }
  StageTimer tm(VISDE_STAGE_K1_PATH_FWD, tcwrec ? 4 : 1, st);
  if (tcwrec) {
    if ((rc = launch_tcw_images(p, p.wimg, true, false, st))) return rc;
    return launch_path_fwd_tcw(p, st);
  }
  if (tcrec) return launch_path_fwd_tc(p, st);
  if (use_fasts(d, p)) return launch_path_fwd_fasts(p, st);
  if (use_fast(d, p)) {
    const int fam = d->variant & 0xff;
    const int nb = (fam == VISDE_VARIANT_FAST || fam == VISDE_VARIANT_TC) ? 0 : tiled_batch_tile(d->B, fam == VISDE_VARIANT_TILED, false);
    return nb > 0 ? launch_path_fwd_tiled(p, nb, st) : launch_path_fwd_fast(p, st);
  }
  return launch_path_fwd_generic(p, st);
}

int visde_path_bwd(const visde_dims* d, float dt, const float* g_paths, const float* g_means,
                   const float* g_chol, const visde_ctx_view* ctx, const float* theta,
                   const float* eps, const visde_weights* w, const float* paths, const void* stash,
                   float* grad_x0, const visde_ctx_grad_view* grad_ctx, float* grad_theta,
                   const visde_weight_grads* gw, void* workspace, size_t workspace_bytes,
                   void* stream) {
  int rc = check_dims(d);
  if (rc) return rc;
  if ((rc = check_weights(d, w))) return rc;
  VISDE_REQUIRE(dt > 0.f, "time_step must be positive");
  VISDE_REQUIRE(gw != nullptr, "weight grads is NULL");
  if (d->B == 0) return VISDE_OK;
  VISDE_REQUIRE(g_paths && grad_x0 && paths, "NULL tensor argument");
  VISDE_REQUIRE(d->T == 0 || (g_means && g_chol && ctx && ctx->ptr && eps && stash && grad_ctx && grad_ctx->ptr),
                "NULL tensor argument");
  VISDE_REQUIRE(d->P == 0 || (theta && grad_theta), "theta / grad_theta is NULL");
  BwdWs ws = bwd_ws(d);
  if (workspace_bytes < ws.total || !workspace) {
    set_error("path_bwd: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return VISDE_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int H = d->H, S = d->S, C = d->C, P = d->P, NL = d->NL, G = 3 * d->H;
  const int ld0 = S + C + P;
  char* wsb = reinterpret_cast<char*>(workspace);
  visde_ctx_view ctx_f32;
  if (d->T > 0 && ctx_convertible(d, ctx)) {
    if ((rc = convert_ctx(d, ctx, reinterpret_cast<float*>(wsb + ws.ctx_f32), &ctx_f32, st))) return rc;
    ctx = &ctx_f32;
  }
  PathParams p;
  fill_common(p, d, dt, w);
  p.x0 = nullptr;
  p.theta = theta;
  p.eps = eps;
  p.g_paths = g_paths;
  p.g_means = g_means;
  p.g_chol = g_chol;
  p.grad_x0 = grad_x0;
  StashLayout sl = stash_layout(d);
  p.stash = const_cast<float*>(reinterpret_cast<const float*>(stash));
  p.raw = reinterpret_cast<float*>(const_cast<char*>(reinterpret_cast<const char*>(stash)) +
                                   align_up(sizeof(float) * sl.h_floats));
  p.dg = reinterpret_cast<float*>(wsb + ws.dg);
  p.dout = reinterpret_cast<float*>(wsb + ws.dout);
  p.sdg = reinterpret_cast<float*>(wsb + ws.sdg);
  p.paths = const_cast<float*>(paths);
  p.otile = reinterpret_cast<float*>(reinterpret_cast<char*>(p.raw) + align_up(sizeof(float) * sl.raw_floats));
  const bool tcwrec = d->T > 0 && use_tcw_rec(d, p, ctx);
  const bool tcrec = (d->T > 0 && use_tc_rec(d, p, ctx)) || tcwrec;
  const bool fastsk = !tcrec && use_fasts(d, p);
  const bool fastk = !tcrec && !fastsk && use_fast(d, p);
  p.cta_part = (fastk || fastsk) ? reinterpret_cast<float*>(wsb + ws.cta_part) : nullptr;
  float* partials = reinterpret_cast<float*>(wsb + ws.partials);

  // K2: reverse-time recurrence
  const float* dg_tiled = nullptr;
  int part_ctas = 0;  // per-CTA partial records written by the batch-tiled backward (0: one record per fast-family CTA)
  if (tcrec) {
    // tensor-core family: reads the tiled stash the forward wrote and emits d_pre / d_out row-fastest tiled; the
    // thin reductions over (b, t) run right behind it, K3 / K4 below read the tiled buffers directly
    StageTimer tm(VISDE_STAGE_K2_PATH_BWD, tcwrec ? 8 : 3, st);
    p.dg = reinterpret_cast<float*>(wsb + ws.dg_tiled);
    dg_tiled = p.dg;
    if (tcwrec) {
      // wide-state family: weight images + tiled cotangent records first, S-sized reductions as time-parallel passes behind
      p.wimg = wsb + ws.wimg;
      p.ctile = reinterpret_cast<float*>(wsb + ws.ctile);
      if ((rc = launch_tcw_images(p, p.wimg, false, true, st))) return rc;
      if ((rc = launch_path_bwd_tcw(p, st))) return rc;
      rc = launch_tcw_thin_grads(p, gw, partials, ws.partial_floats, st);
    } else {
      rc = launch_path_bwd_tc(p, st);
      if (rc) return rc;
      rc = launch_tc_thin_grads(p, p.dout, gw, partials, ws.partial_floats, st);
    }
    if (rc) return rc;
  } else {
    StageTimer tm(VISDE_STAGE_K2_PATH_BWD, fastsk ? 2 : 1, st);
    if (fastk) {
      // more trajectories than SMs: the batch-tiled kernel carries 4-8 trajectories through every barrier / shuffle round
      const int fam = d->variant & 0xff;
      const int nb = (fam == VISDE_VARIANT_FAST || fam == VISDE_VARIANT_TC || p.H > 64 || d->T == 0)
                         ? 0 : tiled_batch_tile(d->B, fam == VISDE_VARIANT_TILED, true);
      rc = nb > 0 ? launch_path_bwd_tiled(p, nb, &part_ctas, st) : launch_path_bwd_fast(p, st);
    } else {
      rc = fastsk ? launch_path_bwd_fasts(p, gw, st) : launch_path_bwd_generic(p, st);
    }
    if (rc) return rc;
  }
  if (d->T == 0) {
    // no steps: every parameter gradient is zero, grad_x0 = g_paths[:, 0]
    for (int k = 0; k < NL; ++k) {
      VISDE_CUDA_CHECK(cudaMemsetAsync(gw->w_ih[k], 0, sizeof(float) * G * (k ? H : ld0), st));
      VISDE_CUDA_CHECK(cudaMemsetAsync(gw->w_hh[k], 0, sizeof(float) * G * H, st));
      VISDE_CUDA_CHECK(cudaMemsetAsync(gw->b_ih[k], 0, sizeof(float) * G, st));
      VISDE_CUDA_CHECK(cudaMemsetAsync(gw->b_hh[k], 0, sizeof(float) * G, st));
    }
    VISDE_CUDA_CHECK(cudaMemsetAsync(gw->out_w, 0, sizeof(float) * p.n_out * H, st));
    VISDE_CUDA_CHECK(cudaMemsetAsync(gw->out_b, 0, sizeof(float) * p.n_out, st));
    if (P) VISDE_CUDA_CHECK(cudaMemsetAsync(grad_theta, 0, sizeof(float) * d->B * P, st));
    return VISDE_OK;
  }

  const int64_t dg_t = (int64_t)NL * kDgSlots * H, dg_b = d->T * dg_t;
  const int64_t st_t = stash_row_floats(NL, H), st_b = d->T * st_t;
  auto dg_src = [&](int k) { return RowSrc{p.dg + (int64_t)k * kDgSlots * H, dg_b, dg_t, 0, kDgSlots * H, VISDE_F32}; };
  auto h_src = [&](int k, int tshift) {
    return RowSrc{p.stash + ((int64_t)k * kStashSlots + kStashH) * H, st_b, st_t, tshift, H, VISDE_F32};
  };
  const RowSrc ones{nullptr, 0, 0, 0, 1, VISDE_F32};

  // the tcgen05 K3 writes fp32 rows with 16-byte stores (needs alignment) or bf16 element-wise
  const bool tc = use_tc(d, ctx) &&
                  (grad_ctx->dtype == VISDE_BF16 ||
                   ((grad_ctx->batch_stride % 4 == 0) && (grad_ctx->time_stride % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(grad_ctx->ptr) & 15) == 0)));
  float* wsplit = reinterpret_cast<float*>(wsb + ws.wsplit);
  // K3: grad_context = d_gi_l0 . W_ih_l0[:, S:S+C]
  {
    StageTimer tm3(VISDE_STAGE_K3_GRAD_CTX, (C > 0) + (P > 0), st);
    if (tc) {
      rc = tc_split_weights(w->w_ih[0], ld0, S, H, C, wsplit, st);
      if (rc) return rc;
      rc = tcrec ? tc_grad_ctx(dg_tiled, dg_t, d->B, d->T, C, H, wsplit, grad_ctx, true, st)
                 : tc_grad_ctx(p.dg, dg_t, d->B, d->T, C, H, wsplit, grad_ctx, false, st);
      if (rc) return rc;
    } else if (C > 0) {
      rc = launch_gemm_nn(dg_src(0), d->B, d->T, G, w->w_ih[0] + S, ld0, C, grad_ctx->ptr, grad_ctx->batch_stride,
                          grad_ctx->time_stride, grad_ctx->dtype, st);
      if (rc) return rc;
    }
    // grad_theta (through the GRU input) = (sum_t d_gi_l0) . W_ih_l0[:, S+C:]
    const bool thin_family = fastk || tcrec || fastsk;  // these families take the theta columns of dW_ih_l0 from sdg
    if (P > 0 && theta_grads_supported(P)) {
      rc = launch_theta_grads(p.sdg, theta, w->w_ih[0], d->B, G, P, ld0, S + C, grad_theta,
                              thin_family ? gw->w_ih[0] : nullptr, st);
      if (rc) return rc;
    } else if (P > 0) {
      RowSrc A{p.sdg, G, 0, 0, G, VISDE_F32};
      rc = launch_gemm_nn(A, d->B, 1, G, w->w_ih[0] + S + C, ld0, P, grad_theta, P, 0, VISDE_F32, st);
      if (rc) return rc;
    }
  }
  // K4: weight gradients
  StageTimer tm4(VISDE_STAGE_K4_WGRAD, (fastk || tcrec || fastsk) ? (tc ? 5 : 3 + 2 * (2 * NL)) : (tc ? 2 + 2 * (2 * NL + 1) : 2 * (2 * NL + 1)), st);
  if (fastk || tcrec || fastsk) {
    // fast family: biases, dW_ih_l0[:, :S], dW_out, db_out were accumulated inside K2 (per-CTA partials);
    // tensor-core family: launch_tc_thin_grads above already wrote them
    if (fastk) {
      rc = launch_fast_partials_reduce(p, gw, st, part_ctas);
      if (rc) return rc;
    }
    if (P > 0 && !theta_grads_supported(P)) {  // theta columns of dW_ih_l0 = (sum_t d_gi_l0)^T theta: only B rows
      RowSrc A{p.sdg, G, 0, 0, G, VISDE_F32};
      RowSrc bs[1] = {RowSrc{theta, P, 0, 0, P, VISDE_F32}};
      TnOut o{gw->w_ih[0] + S + C, ld0, 0, P};
      rc = launch_gemm_tn(A, G, G, 0, bs, 1, d->B, 1, &o, 1, partials, ws.partial_floats, st);
      if (rc) return rc;
    }
    if (fastsk && d->T > 0) {
      // wide-state family: the S-sized reductions (dW_ih_l0[:, :S], dW_out, db_out) are one time-parallel pass over (b, t)
      p.paths = const_cast<float*>(paths);
      rc = launch_fasts_thin_grads(p, gw, partials, ws.partial_floats, st);
      if (rc) return rc;
    }
    if (tcrec) return tc_wgrads_tiled(ctx, dg_tiled, p.stash, d->B, d->T, S, C, P, H, NL, gw, partials, ws.partial_floats, st);
    if (tc) return tc_wgrads(ctx, p.dg, p.stash, d->B, d->T, S, C, P, H, NL, gw, partials, ws.partial_floats, st);
    if (C > 0) {
      RowSrc bs[1] = {RowSrc{ctx->ptr, ctx->batch_stride, ctx->time_stride, 0, C, ctx->dtype}};
      TnOut o{gw->w_ih[0] + S, ld0, 0, C};
      rc = launch_gemm_tn(dg_src(0), G, G, 0, bs, 1, d->B, d->T, &o, 1, partials, ws.partial_floats, st);
      if (rc) return rc;
    }
    for (int k = 0; k < NL; ++k) {
      RowSrc bh[1] = {h_src(k, -1)};
      TnOut oh{gw->w_hh[k], H, 0, H};
      rc = launch_gemm_tn(dg_src(k), G, 2 * H, H, bh, 1, d->B, d->T, &oh, 1, partials, ws.partial_floats, st);
      if (rc) return rc;
      if (k > 0) {
        RowSrc bi[1] = {h_src(k - 1, 0)};
        TnOut oi{gw->w_ih[k], H, 0, H};
        rc = launch_gemm_tn(dg_src(k), G, G, 0, bi, 1, d->B, d->T, &oi, 1, partials, ws.partial_floats, st);
        if (rc) return rc;
      }
    }
    return VISDE_OK;
  }
  if (tc) {
    // tensor-core part: dW_ih0[:, S:S+C], dW_hh_k, dW_ih_1 (tc_gemm.cu); the thin remainder below
    rc = tc_wgrads(ctx, p.dg, p.stash, d->B, d->T, S, C, P, H, NL, gw, partials, ws.partial_floats, st);
    if (rc) return rc;
  }
  {
    RowSrc bs[4];
    int n = 0;
    bs[n++] = RowSrc{paths, (int64_t)(d->T + 1) * S, S, 0, S, VISDE_F32};
    if (C > 0 && !tc) bs[n++] = RowSrc{ctx->ptr, ctx->batch_stride, ctx->time_stride, 0, C, ctx->dtype};
    if (P > 0) bs[n++] = RowSrc{theta, P, 0, 0, P, VISDE_F32};
    bs[n++] = ones;
    if (tc) {
      TnOut outs[3] = {{gw->w_ih[0], ld0, 0, S}, {P > 0 ? gw->w_ih[0] + S + C : nullptr, ld0, S, P}, {gw->b_ih[0], 1, S + P, 1}};
      rc = launch_gemm_tn(dg_src(0), G, G, 0, bs, n, d->B, d->T, outs, 3, partials, ws.partial_floats, st);
    } else {
      TnOut outs[2] = {{gw->w_ih[0], ld0, 0, ld0}, {gw->b_ih[0], 1, ld0, 1}};
      rc = launch_gemm_tn(dg_src(0), G, G, 0, bs, n, d->B, d->T, outs, 2, partials, ws.partial_floats, st);
    }
    if (rc) return rc;
  }
  for (int k = 0; k < NL; ++k) {
    {
      RowSrc bs[2] = {h_src(k, -1), ones};
      TnOut outs[2] = {{gw->w_hh[k], H, 0, H}, {gw->b_hh[k], 1, H, 1}};
      TnOut bias_only{gw->b_hh[k], 1, 0, 1};
      if (tc)
        rc = launch_gemm_tn(dg_src(k), G, 2 * H, H, &bs[1], 1, d->B, d->T, &bias_only, 1, partials, ws.partial_floats, st);
      else
        rc = launch_gemm_tn(dg_src(k), G, 2 * H, H, bs, 2, d->B, d->T, outs, 2, partials, ws.partial_floats, st);
      if (rc) return rc;
    }
    if (k > 0) {
      RowSrc bs[2] = {h_src(k - 1, 0), ones};
      TnOut outs[2] = {{gw->w_ih[k], H, 0, H}, {gw->b_ih[k], 1, H, 1}};
      TnOut bias_only{gw->b_ih[k], 1, 0, 1};
      if (tc)
        rc = launch_gemm_tn(dg_src(k), G, G, 0, &bs[1], 1, d->B, d->T, &bias_only, 1, partials, ws.partial_floats, st);
      else
        rc = launch_gemm_tn(dg_src(k), G, G, 0, bs, 2, d->B, d->T, outs, 2, partials, ws.partial_floats, st);
      if (rc) return rc;
    }
  }
  {
    RowSrc A{p.dout, d->T * (int64_t)p.n_out, p.n_out, 0, p.n_out, VISDE_F32};
    RowSrc bs[2] = {h_src(NL - 1, 0), ones};
    TnOut outs[2] = {{gw->out_w, H, 0, H}, {gw->out_b, 1, H, 1}};
    rc = launch_gemm_tn(A, p.n_out, p.n_out, 0, bs, 2, d->B, d->T, outs, 2, partials, ws.partial_floats, st);
    if (rc) return rc;
  }
  return VISDE_OK;
}

static int fill_elbo(ElboParams& e, const visde_dims* d, float dt, int sde_kind, uint32_t positive_mask,
                     const float* z, const float* means, const float* chol, const float* theta,
                     const float* drift, const float* diffusion, const visde_obs* obs) {
  int rc = check_dims(d);
  if (rc) return rc;
  VISDE_REQUIRE(dt > 0.f, "time_step must be positive");
  VISDE_REQUIRE(sde_kind == VISDE_SDE_GENERIC || sde_kind == VISDE_SDE_OU || sde_kind == VISDE_SDE_LV,
                "unknown sde_kind %d", sde_kind);
  VISDE_REQUIRE(sde_kind != VISDE_SDE_OU || (d->S == 1 && d->P == 3), "OU requires state_dim 1, sde_param_dim 3");
  VISDE_REQUIRE(sde_kind != VISDE_SDE_LV || (d->S == 2 && d->P == 3), "LV requires state_dim 2, sde_param_dim 3");
  VISDE_REQUIRE(d->B == 0 || z, "z is NULL");
  VISDE_REQUIRE(d->B == 0 || d->T == 0 || (means && chol), "means / chol is NULL");
  VISDE_REQUIRE(sde_kind == VISDE_SDE_GENERIC || d->B == 0 || theta, "theta is NULL");
  VISDE_REQUIRE(sde_kind != VISDE_SDE_GENERIC || d->B == 0 || d->T == 0 || (drift && diffusion),
                "generic SDE needs drift and diffusion tensors");
  memset(&e, 0, sizeof(e));
  e.B = d->B;
  e.T = d->T;
  e.S = d->S;
  e.P = d->P;
  e.sde_kind = sde_kind;
  e.pos_mask = positive_mask;
  e.dt = dt;
  e.z = z;
  e.means = means;
  e.chol = chol;
  e.theta = theta;
  e.drift = drift;
  e.diffusion = diffusion;
  if (obs) {
    VISDE_REQUIRE(obs->n_obs >= 0 && obs->obs_dim >= 0, "bad observation sizes");
    VISDE_REQUIRE(obs->n_obs == 0 || (obs->idx && obs->values), "observation pointers are NULL");
    VISDE_REQUIRE(obs->n_obs == 0 || obs->variance > 0.f, "variance must be positive");  // observations.py:47-50
    VISDE_REQUIRE(obs->n_obs == 0 || obs->obs_matrix || obs->obs_dim == d->S,
                  "observation dim %d does not match state dim %d", obs->obs_dim, d->S);
    e.obs = *obs;
  }
  return VISDE_OK;
}

int visde_elbo_fwd(const visde_dims* d, float dt, int sde_kind, uint32_t positive_mask,
                   const float* z, const float* means, const float* chol, const float* theta,
                   const float* drift, const float* diffusion, const visde_obs* obs, float* terms,
                   void* stream) {
  ElboParams e;
  int rc = fill_elbo(e, d, dt, sde_kind, positive_mask, z, means, chol, theta, drift, diffusion, obs);
  if (rc) return rc;
  VISDE_REQUIRE(d->B == 0 || terms, "terms is NULL");
  e.terms = terms;
  StageTimer tm(VISDE_STAGE_K5_ELBO_FWD, 1, (cudaStream_t)stream);
  return launch_elbo_fwd(e, (cudaStream_t)stream);
}

int visde_elbo_bwd(const visde_dims* d, float dt, int sde_kind, uint32_t positive_mask,
                   const float* z, const float* means, const float* chol, const float* theta,
                   const float* drift, const float* diffusion, const visde_obs* obs,
                   const float* g_terms, float* g_z, float* g_means, float* g_chol,
                   float* g_theta, float* g_drift, float* g_diffusion, void* stream) {
  ElboParams e;
  int rc = fill_elbo(e, d, dt, sde_kind, positive_mask, z, means, chol, theta, drift, diffusion, obs);
  if (rc) return rc;
  VISDE_REQUIRE(d->B == 0 || (g_terms && g_z), "g_terms / g_z is NULL");
  VISDE_REQUIRE(d->B == 0 || d->T == 0 || (g_means && g_chol), "g_means / g_chol is NULL");
  VISDE_REQUIRE(d->B == 0 || d->P == 0 || g_theta, "g_theta is NULL");
  VISDE_REQUIRE(sde_kind != VISDE_SDE_GENERIC || d->B == 0 || d->T == 0 || (g_drift && g_diffusion),
                "generic SDE needs g_drift and g_diffusion");
  e.g_terms = g_terms;
  e.g_z = g_z;
  e.g_means = g_means;
  e.g_chol = g_chol;
  e.g_theta = g_theta;
  e.g_drift = g_drift;
  e.g_diffusion = g_diffusion;
  StageTimer tm(VISDE_STAGE_K6_ELBO_BWD, 1, (cudaStream_t)stream);
  return launch_elbo_bwd(e, (cudaStream_t)stream);
}

int visde_profile_begin(int max_records) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  VISDE_REQUIRE(!g_prof_on, "profiler already enabled");
  VISDE_REQUIRE(max_records > 0 && max_records <= (1 << 20), "bad max_records");
  g_prof.resize(max_records);
  for (auto& r : g_prof) {
    VISDE_CUDA_CHECK(cudaEventCreate(&r.beg));
    VISDE_CUDA_CHECK(cudaEventCreate(&r.end));
  }
  g_prof_used = 0;
  g_prof_on = true;
  return VISDE_OK;
}

int visde_profile_end(double* ms_per_stage, int* launches_per_stage) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  VISDE_REQUIRE(g_prof_on, "profiler not enabled");
  g_prof_on = false;
  for (int s = 0; s < VISDE_NUM_STAGES; ++s) {
    if (ms_per_stage) ms_per_stage[s] = 0.0;
    if (launches_per_stage) launches_per_stage[s] = 0;
  }
  int rc = VISDE_OK;
  for (size_t i = 0; i < g_prof_used; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(g_prof[i].end) != cudaSuccess || cudaEventElapsedTime(&ms, g_prof[i].beg, g_prof[i].end) != cudaSuccess) {
      set_error("profiler: event query failed");
      rc = VISDE_ECUDA;
      continue;
    }
    if (ms_per_stage) ms_per_stage[g_prof[i].stage] += ms;
    if (launches_per_stage) launches_per_stage[g_prof[i].stage] += g_prof[i].launches;
  }
  for (auto& r : g_prof) {
    cudaEventDestroy(r.beg);
    cudaEventDestroy(r.end);
  }
  g_prof.clear();
  g_prof_used = 0;
  return rc;
}

// ------------------------------------------------------------------------------------------
// host-buffer session
// ------------------------------------------------------------------------------------------
// Two input sets so that the H2D copies of iteration i+1 (copy stream) overlap the kernels of
// iteration i (compute stream): visde_session_submit / visde_session_wait.  Everything downstream
// of the inputs (paths, stash, workspaces, gradients) exists once: the compute stream serialises it.
struct visde_session_inputs {
  float *x0, *theta, *eps;
  void* ctx;  // [B,T+1,C] in the session's context dtype
  visde_weights w;
  int32_t* obs_idx;
  float *obs_values, *obs_matrix;
  cudaEvent_t loaded;    // copy stream: this set's H2D copies are complete
  cudaEvent_t consumed;  // compute stream: the kernels that read this set are complete
  cudaEvent_t done;      // compute stream: kernels + D2H of the iteration that used this set
  // the kernel sequence of one iteration on this input set, captured once as a CUDA graph (one launch per iteration,
  // no gaps between the dependent kernels); re-captured when a scalar baked into the kernel parameters changes
  cudaGraphExec_t graph[3];  // [0]: the whole iteration, or (user-SDE sessions) one graph per segment
  float graph_dt, graph_var;
  bool graph_obsmat, graph_failed;
  uint32_t uses;
};

struct visde_session {
  visde_dims d;
  int sde_kind;
  uint32_t pos_mask;
  int n_obs, obs_dim;
  int ctx_dtype;          // VISDE_F32 | VISDE_BF16: host context and grad_context element type
  visde_user_sde user;    // sde_kind GENERIC: caller-evaluated drift / diffusion hooks
  cudaStream_t st, copy_st;
  std::vector<void*> allocs;
  visde_session_inputs in[2];
  uint64_t n_submitted, n_waited;
  uint64_t noise_seed;    // eps == NULL: iteration i draws Philox standard normals with key noise_seed + i on the device
  size_t eps_bytes;       // H2D bytes of eps (not copied by iterations that draw on the device)
  bool device_noise;      // the last submitted iteration drew its noise on the device
  // device buffers
  float *paths, *means, *chol, *terms, *g_terms;
  float *g_z, *g_means, *g_chol, *g_theta_elbo, *grad_x0, *grad_theta;
  void* grad_ctx;  // [B,T+1,C] in the context dtype
  float *x_state, *drift, *diffusion, *g_drift, *g_diffusion, *g_x, *g_theta_sde;  // user-SDE sessions only
  float* ctx_f32;  // bf16 sessions: the context widened ONCE per iteration (forward and backward both read it)
  void *stash, *ws_f, *ws_b;
  size_t ws_f_bytes, ws_b_bytes;
  visde_weight_grads gw;
  size_t h2d, d2h;
  int launches;
};

}  // extern "C"

namespace {
template <typename Tp>
int dev_alloc(visde_session* s, Tp** p, size_t bytes) {
  void* q = nullptr;
  VISDE_CUDA_CHECK(cudaMalloc(&q, bytes ? bytes : 16));
  s->allocs.push_back(q);
  *p = reinterpret_cast<Tp*>(q);
  return VISDE_OK;
}
size_t w_ih_floats(const visde_dims& d, int k) { return (size_t)3 * d.H * (k ? d.H : d.S + d.C + d.P); }
size_t ctx_elem_bytes(const visde_session* s) { return s->ctx_dtype == VISDE_BF16 ? 2 : 4; }

int session_alloc(visde_session* s) {
  const visde_dims* d = &s->d;
  const size_t B = d->B, T = d->T, S = d->S, C = d->C, P = d->P, H = d->H, G = 3 * H;
  const size_t n_out = S + S * (S + 1) / 2;
  int rc;
#define A_(ptr, n) if ((rc = dev_alloc(s, &(ptr), sizeof(float) * (n)))) return rc;
  for (int q = 0; q < 2; ++q) {
    visde_session_inputs& in = s->in[q];
    A_(in.x0, B * S) A_(in.theta, B * P) A_(in.eps, B * T * S)
    if ((rc = dev_alloc(s, &in.ctx, ctx_elem_bytes(s) * B * (T + 1) * C))) return rc;
    A_(in.obs_values, (size_t)s->n_obs * s->obs_dim) A_(in.obs_matrix, (size_t)s->obs_dim * S)
    if ((rc = dev_alloc(s, &in.obs_idx, sizeof(int32_t) * s->n_obs))) return rc;
    for (int k = 0; k < d->NL; ++k) {
      float *a, *b2, *c2, *e2;
      A_(a, w_ih_floats(*d, k)) A_(b2, G * H) A_(c2, G) A_(e2, G)
      in.w.w_ih[k] = a; in.w.w_hh[k] = b2; in.w.b_ih[k] = c2; in.w.b_hh[k] = e2;
    }
    float *a, *b2;
    A_(a, n_out * H) A_(b2, n_out)
    in.w.out_w = a; in.w.out_b = b2;
    VISDE_CUDA_CHECK(cudaEventCreateWithFlags(&in.loaded, cudaEventDisableTiming));
    VISDE_CUDA_CHECK(cudaEventCreateWithFlags(&in.consumed, cudaEventDisableTiming));
    VISDE_CUDA_CHECK(cudaEventCreateWithFlags(&in.done, cudaEventDisableTiming));
  }
  A_(s->paths, B * (T + 1) * S) A_(s->means, B * T * S) A_(s->chol, B * T * S * S) A_(s->terms, B * 4) A_(s->g_terms, B * 4)
  A_(s->g_z, B * (T + 1) * S) A_(s->g_means, B * T * S) A_(s->g_chol, B * T * S * S) A_(s->g_theta_elbo, B * P)
  A_(s->grad_x0, B * S) A_(s->grad_theta, B * P)
  if ((rc = dev_alloc(s, &s->grad_ctx, ctx_elem_bytes(s) * B * (T + 1) * C))) return rc;
  {
    const visde_ctx_view host_layout{s->in[0].ctx, (int64_t)((T + 1) * C), (int64_t)C, VISDE_BF16};
    if (s->ctx_dtype == VISDE_BF16 && ctx_convertible(d, &host_layout)) A_(s->ctx_f32, B * T * C)
  }
  if (s->sde_kind == VISDE_SDE_GENERIC) {
    A_(s->x_state, B * T * S) A_(s->drift, B * T * S) A_(s->diffusion, B * T * S * S) A_(s->g_drift, B * T * S)
    A_(s->g_diffusion, B * T * S * S) A_(s->g_x, B * T * S) A_(s->g_theta_sde, B * P)
  }
  if ((rc = dev_alloc(s, &s->stash, visde_stash_bytes(d)))) return rc;
  s->ws_f_bytes = visde_workspace_bytes(d, 0);
  s->ws_b_bytes = visde_workspace_bytes(d, 1);
  if ((rc = dev_alloc(s, &s->ws_f, s->ws_f_bytes))) return rc;
  if ((rc = dev_alloc(s, &s->ws_b, s->ws_b_bytes))) return rc;
  for (int k = 0; k < d->NL; ++k) {
    float *ga, *gb, *gc, *ge;
    A_(ga, w_ih_floats(*d, k)) A_(gb, G * H) A_(gc, G) A_(ge, G)
    s->gw.w_ih[k] = ga; s->gw.w_hh[k] = gb; s->gw.b_ih[k] = gc; s->gw.b_hh[k] = ge;
  }
  float *ga, *gb;
  A_(ga, n_out * H) A_(gb, n_out)
  s->gw.out_w = ga; s->gw.out_b = gb;
#undef A_
  return VISDE_OK;
}
}  // namespace

extern "C" {

int visde_session_create(const visde_dims* d, int sde_kind, uint32_t positive_mask, int32_t n_obs,
                         int32_t obs_dim, int32_t ctx_dtype, const visde_user_sde* user_sde, visde_session** out) {
  int rc = check_dims(d);
  if (rc) return rc;
  VISDE_REQUIRE(out != nullptr, "out is NULL");
  VISDE_REQUIRE(sde_kind == VISDE_SDE_OU || sde_kind == VISDE_SDE_LV || sde_kind == VISDE_SDE_GENERIC,
                "unknown sde_kind %d", sde_kind);
  VISDE_REQUIRE(sde_kind != VISDE_SDE_GENERIC || (user_sde && user_sde->eval && user_sde->vjp),
                "a user-SDE session needs the eval and vjp hooks (visde_user_sde)");
  VISDE_REQUIRE(ctx_dtype == VISDE_F32 || ctx_dtype == VISDE_BF16, "context dtype must be VISDE_F32 or VISDE_BF16");
  VISDE_REQUIRE(d->B > 0 && d->T > 0, "session needs B > 0 and T > 0");
  VISDE_REQUIRE(n_obs >= 0 && obs_dim >= 0, "bad observation sizes");
  visde_session* s = new visde_session();
  s->d = *d;
  s->sde_kind = sde_kind;
  s->pos_mask = positive_mask;
  s->n_obs = n_obs;
  s->obs_dim = obs_dim;
  s->ctx_dtype = ctx_dtype;
  if (user_sde) s->user = *user_sde;
  const size_t B = d->B, T = d->T, S = d->S, C = d->C, P = d->P, H = d->H, G = 3 * H;
  const size_t n_out = S + S * (S + 1) / 2;
  if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&s->copy_st, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    visde_session_destroy(s);
    return VISDE_ECUDA;
  }
  if ((rc = session_alloc(s))) {
    visde_session_destroy(s);
    return rc;
  }
  size_t wfloats = n_out * H + n_out;
  for (int k = 0; k < d->NL; ++k) wfloats += w_ih_floats(*d, k) + G * H + 2 * G;
  // grad_ctx row T is never written by the kernels: zero it once
  cudaMemsetAsync(s->grad_ctx, 0, ctx_elem_bytes(s) * B * (T + 1) * C, s->st);
  cudaStreamSynchronize(s->st);
  s->h2d = sizeof(float) * (B * S + B * P + B * T * S + wfloats + (size_t)n_obs * obs_dim) +
           ctx_elem_bytes(s) * B * (T + 1) * C + sizeof(int32_t) * n_obs;
  s->eps_bytes = sizeof(float) * B * T * S;
  s->noise_seed = 0;
  s->device_noise = false;
  s->d2h = sizeof(float) * (B * 4 + B * S + B * P + wfloats);
  // K0, K1, elbo fwd, cotangent fill, elbo bwd, K2, K3, grad_theta gemm, (2 NL + 1) x (tn + reduce), add
  s->launches = 8 + 2 * (2 * d->NL + 1) + 1;
  *out = s;
  return VISDE_OK;
}

void visde_session_destroy(visde_session* s) {
  if (!s) return;
  if (s->st) cudaStreamSynchronize(s->st);
  if (s->copy_st) cudaStreamSynchronize(s->copy_st);
  for (void* p : s->allocs) cudaFree(p);
  for (int q = 0; q < 2; ++q) {
    if (s->in[q].loaded) cudaEventDestroy(s->in[q].loaded);
    if (s->in[q].consumed) cudaEventDestroy(s->in[q].consumed);
    if (s->in[q].done) cudaEventDestroy(s->in[q].done);
    for (int g = 0; g < 3; ++g)
      if (s->in[q].graph[g]) cudaGraphExecDestroy(s->in[q].graph[g]);
  }
  if (s->st) cudaStreamDestroy(s->st);
  if (s->copy_st) cudaStreamDestroy(s->copy_st);
  delete s;
}

size_t visde_session_h2d_bytes(const visde_session* s) { return s ? s->h2d - (s->device_noise ? s->eps_bytes : 0) : 0; }
int visde_session_set_noise_seed(visde_session* s, uint64_t seed) {
  VISDE_REQUIRE(s != nullptr, "session_set_noise_seed: NULL session");
  s->noise_seed = seed;
  return VISDE_OK;
}
size_t visde_session_d2h_bytes(const visde_session* s) { return s ? s->d2h : 0; }
int visde_session_launches(const visde_session* s) { return s ? s->launches : 0; }

// kernel sequence of one iteration on input set `in` (everything between the H2D and the D2H copies), in three segments
// around the two places where a user-SDE session hands control to the caller's hooks:
//   segment 0: context widening, path forward, x_t = to_state(z_t)      -> [hook: drift / diffusion]
//   segment 1: ELBO forward, loss cotangent, ELBO backward              -> [hook: their vector-Jacobian product]
//   segment 2: chain into g_z, path backward, grad_theta sums
static int session_segment(visde_session* s, visde_session_inputs& in, float dt, const visde_obs* od_p, cudaStream_t st, int seg) {
  const visde_dims& d = s->d;
  const size_t B = d.B, T = d.T, C = d.C, P = d.P;
  const visde_obs& od = *od_p;
  visde_ctx_view cv{in.ctx, (int64_t)((T + 1) * C), (int64_t)C, s->ctx_dtype};
  visde_ctx_grad_view gv{s->grad_ctx, (int64_t)((T + 1) * C), (int64_t)C, s->ctx_dtype};
  const bool generic = s->sde_kind == VISDE_SDE_GENERIC;
  const int64_t n_x = (int64_t)B * T * d.S;
  int rc;
  if (s->ctx_f32) {  // bf16 host context: widen once, both directions consume the fp32 copy (grad_context stays bf16)
    visde_ctx_view wide{s->ctx_f32, (int64_t)(T * C), (int64_t)C, VISDE_F32};
    if (seg == 0 && (rc = convert_ctx(&d, &cv, s->ctx_f32, &wide, st))) return rc;
    cv = wide;
  }
  if (seg == 0) {
    rc = visde_path_fwd(&d, dt, in.x0, &cv, in.theta, in.eps, &in.w, s->paths, s->means, s->chol, s->stash, s->ws_f,
                        s->ws_f_bytes, st);
    if (rc) return rc;
    if (generic) {
      // evidence_lower_bound.py:31-40: x_t = to_state(z_t), then the caller's drift / diffusion on the [B*T, S] rows
      state_rows_kernel<<<(unsigned)((n_x + 255) / 256), 256, 0, st>>>(s->paths, d.B, d.T, d.S, s->pos_mask, s->x_state);
      VISDE_CUDA_CHECK(cudaGetLastError());
    }
    return VISDE_OK;
  }
  if (seg == 1) {
    rc = visde_elbo_fwd(&d, dt, s->sde_kind, s->pos_mask, s->paths, s->means, s->chol, in.theta, s->drift, s->diffusion, &od,
                        s->terms, st);
    if (rc) return rc;
    fill_loss_cotangent_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(s->g_terms, (int64_t)B);
    VISDE_CUDA_CHECK(cudaGetLastError());
    return visde_elbo_bwd(&d, dt, s->sde_kind, s->pos_mask, s->paths, s->means, s->chol, in.theta, s->drift, s->diffusion, &od,
                          s->g_terms, s->g_z, s->g_means, s->g_chol, s->g_theta_elbo, s->g_drift, s->g_diffusion, st);
  }
  if (generic) {
    chain_state_grad_kernel<<<(unsigned)((n_x + 255) / 256), 256, 0, st>>>(s->paths, s->g_x, d.B, d.T, d.S, s->pos_mask, s->g_z);
    VISDE_CUDA_CHECK(cudaGetLastError());
  }
  rc = visde_path_bwd(&d, dt, s->g_z, s->g_means, s->g_chol, &cv, in.theta, in.eps, &in.w, s->paths, s->stash,
                      s->grad_x0, &gv, s->grad_theta, &s->gw, s->ws_b, s->ws_b_bytes, st);
  if (rc) return rc;
  add_inplace_kernel<<<(unsigned)((B * P + 255) / 256), 256, 0, st>>>(s->grad_theta, s->g_theta_elbo, (int64_t)(B * P));
  VISDE_CUDA_CHECK(cudaGetLastError());
  if (generic) {
    add_inplace_kernel<<<(unsigned)((B * P + 255) / 256), 256, 0, st>>>(s->grad_theta, s->g_theta_sde, (int64_t)(B * P));
    VISDE_CUDA_CHECK(cudaGetLastError());
  }
  return VISDE_OK;
}

// the caller's hook after segment `seg` of a user-SDE session (launched on the compute stream, outside any capture)
static int session_hook(visde_session* s, visde_session_inputs& in, cudaStream_t st, int seg) {
  if (s->sde_kind != VISDE_SDE_GENERIC) return VISDE_OK;
  if (seg == 0 && s->user.eval(s->user.user, s->x_state, in.theta, s->drift, s->diffusion, st) != 0) {
    set_error("session: the user SDE eval hook failed");
    return VISDE_EINVAL;
  }
  if (seg == 1 && s->user.vjp(s->user.user, s->x_state, in.theta, s->g_drift, s->g_diffusion, s->g_x, s->g_theta_sde, st) != 0) {
    set_error("session: the user SDE vjp hook failed");
    return VISDE_EINVAL;
  }
  return VISDE_OK;
}

// segments [seg0, seg1) launched kernel by kernel, the hooks between / behind them included
static int session_enqueue(visde_session* s, visde_session_inputs& in, float dt, const visde_obs* od_p, cudaStream_t st,
                           int seg0 = 0, int seg1 = 3) {
  for (int seg = seg0; seg < seg1; ++seg) {
    int rc = session_segment(s, in, dt, od_p, st, seg);
    if (!rc) rc = session_hook(s, in, st, seg);
    if (rc) return rc;
  }
  return VISDE_OK;
}

// capture segments [seg0, seg1) (no hook inside) into an executable graph; nullptr when capture is unavailable
static cudaGraphExec_t session_capture(visde_session* s, visde_session_inputs& in, float dt, const visde_obs* od_p, cudaStream_t st,
                                       int seg0, int seg1, int* rc_out) {
  cudaGraph_t g = nullptr;
  cudaGraphExec_t ge = nullptr;
  *rc_out = VISDE_OK;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  int rc = VISDE_OK;
  for (int seg = seg0; seg < seg1 && !rc; ++seg) rc = session_segment(s, in, dt, od_p, st, seg);
  const cudaError_t ce = cudaStreamEndCapture(st, &g);
  if (rc || ce != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    return nullptr;  // the caller launches the segment plainly (and reports its error, if it has one)
  }
  if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) {
    ge = nullptr;
    cudaGetLastError();
  }
  cudaGraphDestroy(g);
  return ge;
}

int visde_session_submit(visde_session* s, float dt, const float* x0, const void* context,
                         const float* theta, const float* eps, const visde_weights* w_host,
                         const visde_obs* obs_host, float* terms, float* grad_x0, float* grad_theta,
                         const visde_weight_grads* gw_host, void* grad_context) {
  VISDE_REQUIRE(s && x0 && context && theta && w_host && obs_host && terms && grad_x0 && grad_theta && gw_host,
                "session_submit: NULL argument");
  VISDE_REQUIRE(obs_host->n_obs == s->n_obs && obs_host->obs_dim == s->obs_dim, "session_submit: observation shape changed");
  VISDE_REQUIRE(s->n_submitted - s->n_waited < 2, "session_submit: two iterations already in flight; call visde_session_wait");
  const visde_dims& d = s->d;
  const size_t B = d.B, T = d.T, S = d.S, C = d.C, P = d.P, H = d.H, G = 3 * H;
  const size_t n_out = S + S * (S + 1) / 2;
  visde_session_inputs& in = s->in[s->n_submitted & 1];
  cudaStream_t st = s->st, cs = s->copy_st;
  // this set was last read by iteration n_submitted - 2
  VISDE_CUDA_CHECK(cudaStreamWaitEvent(cs, in.consumed, 0));
#define H2D(dst, src, n) VISDE_CUDA_CHECK(cudaMemcpyAsync((void*)(dst), (src), sizeof(float) * (n), cudaMemcpyHostToDevice, cs))
#define D2H(dst, src, n) VISDE_CUDA_CHECK(cudaMemcpyAsync((dst), (src), sizeof(float) * (n), cudaMemcpyDeviceToHost, st))
  H2D(in.x0, x0, B * S);
  VISDE_CUDA_CHECK(cudaMemcpyAsync(in.ctx, context, ctx_elem_bytes(s) * B * (T + 1) * C, cudaMemcpyHostToDevice, cs));
  H2D(in.theta, theta, B * P);
  if (eps) H2D(in.eps, eps, B * T * S);
  s->device_noise = eps == nullptr;
  for (int k = 0; k < d.NL; ++k) {
    H2D(in.w.w_ih[k], w_host->w_ih[k], w_ih_floats(d, k));
    H2D(in.w.w_hh[k], w_host->w_hh[k], G * H);
    H2D(in.w.b_ih[k], w_host->b_ih[k], G);
    H2D(in.w.b_hh[k], w_host->b_hh[k], G);
  }
  H2D(in.w.out_w, w_host->out_w, n_out * H);
  H2D(in.w.out_b, w_host->out_b, n_out);
  visde_obs od = *obs_host;
  if (s->n_obs) {
    H2D(in.obs_values, obs_host->values, (size_t)s->n_obs * s->obs_dim);
    VISDE_CUDA_CHECK(cudaMemcpyAsync(in.obs_idx, obs_host->idx, sizeof(int32_t) * s->n_obs, cudaMemcpyHostToDevice, cs));
    if (obs_host->obs_matrix) H2D(in.obs_matrix, obs_host->obs_matrix, (size_t)s->obs_dim * S);
  }
  od.idx = in.obs_idx;
  od.values = in.obs_values;
  od.obs_matrix = obs_host->obs_matrix ? in.obs_matrix : nullptr;
  VISDE_CUDA_CHECK(cudaEventRecord(in.loaded, cs));
  VISDE_CUDA_CHECK(cudaStreamWaitEvent(st, in.loaded, 0));
  if (!eps) {
    // diffusion_path_sampler.py:57 draws the noise on the device: here the counter-based stream of visde_philox_normal, a
    // fresh key per iteration (ahead of the graph replay: the key is a kernel argument)
    int prc = visde_philox_normal(s->noise_seed + s->n_submitted, d.B, d.T, d.S, in.eps, st);
    if (prc) return prc;
  }

  // first use of an input set runs kernel by kernel (per-kernel attributes, tensor-map encoders); from the second use
  // on the same sequence is replayed from CUDA graphs: ONE graph per iteration, or -- user-SDE sessions, whose hooks run
  // between the segments on the same stream -- one graph per segment
  int rc;
  const bool obs_mat = obs_host->obs_matrix != nullptr;
  const bool generic = s->sde_kind == VISDE_SDE_GENERIC;
  const int nseg = generic ? 3 : 1;
  if (in.graph[0] && (in.graph_dt != dt || in.graph_var != od.variance || in.graph_obsmat != obs_mat)) {
    for (int g = 0; g < 3; ++g) {
      if (in.graph[g]) cudaGraphExecDestroy(in.graph[g]);
      in.graph[g] = nullptr;
    }
  }
  const bool use_graphs = in.uses >= 1 && !g_prof_on;
  if (use_graphs && !in.graph[0] && !in.graph_failed) {
    in.graph_dt = dt;
    in.graph_var = od.variance;
    in.graph_obsmat = obs_mat;
  }
  for (int g = 0; g < nseg; ++g) {
    const int seg0 = generic ? g : 0, seg1 = generic ? g + 1 : 3;
    if (use_graphs && !in.graph[g] && !in.graph_failed) {
      in.graph[g] = session_capture(s, in, dt, &od, st, seg0, seg1, &rc);
      if (!in.graph[g]) in.graph_failed = true;  // capture unavailable: plain launches from here on
    }
    if (use_graphs && in.graph[g]) {
      VISDE_CUDA_CHECK(cudaGraphLaunch(in.graph[g], st));
      if ((rc = session_hook(s, in, st, seg1 - 1))) return rc;
    } else if ((rc = session_enqueue(s, in, dt, &od, st, seg0, seg1))) {
      return rc;
    }
  }
  ++in.uses;
  VISDE_CUDA_CHECK(cudaEventRecord(in.consumed, st));

  D2H(terms, s->terms, B * 4);
  D2H(grad_x0, s->grad_x0, B * S);
  D2H(grad_theta, s->grad_theta, B * P);
  for (int k = 0; k < d.NL; ++k) {
    D2H(gw_host->w_ih[k], s->gw.w_ih[k], w_ih_floats(d, k));
    D2H(gw_host->w_hh[k], s->gw.w_hh[k], G * H);
    D2H(gw_host->b_ih[k], s->gw.b_ih[k], G);
    D2H(gw_host->b_hh[k], s->gw.b_hh[k], G);
  }
  D2H(gw_host->out_w, s->gw.out_w, n_out * H);
  D2H(gw_host->out_b, s->gw.out_b, n_out);
  if (grad_context)
    VISDE_CUDA_CHECK(cudaMemcpyAsync(grad_context, s->grad_ctx, ctx_elem_bytes(s) * B * (T + 1) * C, cudaMemcpyDeviceToHost, st));
#undef H2D
#undef D2H
  VISDE_CUDA_CHECK(cudaEventRecord(in.done, st));
  ++s->n_submitted;
  return VISDE_OK;
}

int visde_session_wait(visde_session* s) {
  VISDE_REQUIRE(s != nullptr, "session_wait: NULL session");
  VISDE_REQUIRE(s->n_waited < s->n_submitted, "session_wait: nothing in flight");
  VISDE_CUDA_CHECK(cudaEventSynchronize(s->in[s->n_waited & 1].done));
  ++s->n_waited;
  return VISDE_OK;
}

int visde_session_step(visde_session* s, float dt, const float* x0, const void* context,
                       const float* theta, const float* eps, const visde_weights* w_host,
                       const visde_obs* obs_host, float* terms, float* grad_x0, float* grad_theta,
                       const visde_weight_grads* gw_host, void* grad_context) {
  VISDE_REQUIRE(s != nullptr, "session_step: NULL session");
  VISDE_REQUIRE(s->n_submitted == s->n_waited, "session_step: iterations still in flight; call visde_session_wait first");
  int rc = visde_session_submit(s, dt, x0, context, theta, eps, w_host, obs_host, terms, grad_x0, grad_theta, gw_host,
                                grad_context);
  if (rc) return rc;
  return visde_session_wait(s);
}

}  // extern "C"
