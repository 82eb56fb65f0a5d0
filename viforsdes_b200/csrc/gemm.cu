// Time-parallel GEMMs around the recurrence (fp32 SIMT, 64x64x16 tiles, 4x4 per thread).
//
//   K0  gemm_nt : gi_ctx[(b,t), 3H] = ctx[b,t,:] . W_ih_l0[:, S:S+C]^T + b_ih_l0   (forward)
//   K3  gemm_nn : grad_ctx[b,t,:]   = d_gi_l0[(b,t), :] . W_ih_l0[:, S:S+C]         (backward)
//   K4  gemm_tn : every weight gradient as sum_{(b,t)} d_pre (x) input, split-K over CTAs with a
//                 fixed-order second-stage reduction (deterministic; replaces the reference's
//                 ~87k global atomics per trajectory-step, kernels/backward.py:108-139,534-590).
// Rows are gathered as (b, t) = (k / T, k % T) with per-source batch/time strides, so the strided
// context[:, :-1] view and bf16 context are consumed in place (no copy, cf. kernels/forward.py:495).
#include "common.cuh"

namespace visde {
namespace {

constexpr int TM = 64, TN = 64, TK = 16, NTH = 256;
constexpr int LDS = TM + 4;  // smem leading dim (keeps float4 alignment: 68*4 B = 272 B)

__device__ __forceinline__ float load_elem(const void* base, int64_t off, int dtype) {
  return dtype == VISDE_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off])
                             : reinterpret_cast<const float*>(base)[off];
}

__device__ __forceinline__ void mma_tile(const float (*As)[LDS], const float (*Bs)[LDS], int ty, int tx,
                                         float (&acc)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < TK; ++kk) {
    float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// ---------------------------------------------------------------------------------------------
// NT / NN: M = B*T gathered rows (k-contiguous A), weights as B operand.
//   B_KCONTIG = true : Bop(k, n) = W[n*ldw + k]   (NT, context projection)
//   B_KCONTIG = false: Bop(k, n) = W[k*ldw + n]   (NN, grad_context)
// ---------------------------------------------------------------------------------------------
struct RowOut {
  void* ptr;
  int64_t bstride, tstride;  // destination row (b,t) at ptr + b*bstride + t*tstride
  int dtype;
  const float* bias;  // [N] or nullptr
};

template <bool B_KCONTIG>
__global__ void __launch_bounds__(NTH) gemm_rows_kernel(RowSrc A, int64_t M, int64_t T, int K,
                                                        const float* __restrict__ W, int ldw, int N,
                                                        RowOut out) {
  __shared__ __align__(16) float As[TK][LDS];
  __shared__ __align__(16) float Bs[TK][LDS];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;

  // A loader: kk = tid % 16, rows mm = tid / 16 + 16 * pass
  int64_t aoff[4];
  bool aok[4];
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    int64_t r = m0 + ty + 16 * ps;
    aok[ps] = r < M;
    int64_t b = aok[ps] ? r / T : 0, t = aok[ps] ? r % T : 0;
    t += A.tshift;
    if (t < 0) aok[ps] = false;
    aoff[ps] = b * A.bstride + t * A.tstride;
  }
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      int k = k0 + tx;
      float v = (aok[ps] && k < K) ? load_elem(A.base, aoff[ps] + k, A.dtype) : 0.f;
      As[tx][ty + 16 * ps] = v;
    }
    if (B_KCONTIG) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        int k = k0 + tx, n = n0 + ty + 16 * ps;
        Bs[tx][ty + 16 * ps] = (k < K && n < N) ? W[(int64_t)n * ldw + k] : 0.f;
      }
    } else {
      const int nn = tid % 64;
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        int kk = tid / 64 + 4 * ps;
        int k = k0 + kk, n = n0 + nn;
        Bs[kk][nn] = (k < K && n < N) ? W[(int64_t)k * ldw + n] : 0.f;
      }
    }
    __syncthreads();
    mma_tile(As, Bs, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = m0 + ty * 4 + i;
    if (r >= M) continue;
    int64_t b = r / T, t = r % T;
    int64_t o = b * out.bstride + t * out.tstride;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (out.bias ? out.bias[n] : 0.f);
      if (out.dtype == VISDE_BF16)
        reinterpret_cast<__nv_bfloat16*>(out.ptr)[o + n] = __float2bfloat16(v);
      else
        reinterpret_cast<float*>(out.ptr)[o + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TN split-K: C[m, n] = sum_k A[k, amap(m)] * Bcat[k, n]
// ---------------------------------------------------------------------------------------------
struct TnArgs {
  RowSrc A;
  int M, a_split, a_skip;
  RowSrc Bsrc[4];
  int nsrc;
  int N;
  int64_t K, T, kchunk;
  float* partials;  // [nsplit][M][N]
};

__global__ void __launch_bounds__(NTH) gemm_tn_kernel(TnArgs g) {
  __shared__ __align__(16) float As[TK][LDS];
  __shared__ __align__(16) float Bs[TK][LDS];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.kchunk;
  const int64_t kend = kbeg + g.kchunk < g.K ? kbeg + g.kchunk : g.K;

  // loaders: column = tid % 64 (contiguous in memory), k rows = tid / 64 + 4 * pass
  const int col = tid % 64, kr = tid / 64;
  const int m = m0 + col;
  const bool m_ok = m < g.M;
  const int acol = m < g.a_split ? m : m + g.a_skip;
  const int n = n0 + col;
  int bsrc = -1, bcol = 0;
  {
    int c = n;
    for (int s = 0; s < g.nsrc; ++s) {
      if (c < g.Bsrc[s].ncols) { bsrc = s; bcol = c; break; }
      c -= g.Bsrc[s].ncols;
    }
  }
  RowSrc B = g.Bsrc[bsrc < 0 ? 0 : bsrc];

  int64_t rb[4], rt[4];
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    int64_t k = kbeg + kr + 4 * ps;
    rb[ps] = k / g.T;
    rt[ps] = k % g.T;
  }
  float acc[4][4] = {};
  for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) {
      const int kk = kr + 4 * ps;
      const bool k_ok = k0 + kk < kend;
      float av = 0.f, bv = 0.f;
      if (k_ok && m_ok) {
        int64_t t = rt[ps] + g.A.tshift;
        if (t >= 0) av = load_elem(g.A.base, rb[ps] * g.A.bstride + t * g.A.tstride + acol, g.A.dtype);
      }
      if (k_ok && bsrc >= 0) {
        if (B.base == nullptr) {
          bv = 1.f;
        } else {
          int64_t t = rt[ps] + B.tshift;
          if (t >= 0) bv = load_elem(B.base, rb[ps] * B.bstride + t * B.tstride + bcol, B.dtype);
        }
      }
      As[kk][col] = av;
      Bs[kk][col] = bv;
      rt[ps] += TK;
      while (rt[ps] >= g.T) { rt[ps] -= g.T; ++rb[ps]; }
    }
    __syncthreads();
    mma_tile(As, Bs, ty, tx, acc);
    __syncthreads();
  }
  float* P = g.partials + (int64_t)blockIdx.z * g.M * g.N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int mm = m0 + ty * 4 + i;
    if (mm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int nn = n0 + tx * 4 + j;
      if (nn < g.N) P[(int64_t)mm * g.N + nn] = acc[i][j];
    }
  }
}

struct TnReduceArgs {
  const float* partials;
  int nsplit, M, N;
  TnOut outs[6];
  int nouts;
};

__global__ void gemm_tn_reduce_kernel(TnReduceArgs r) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= r.M * r.N) return;
  int m = idx / r.N, n = idx % r.N;
  float acc = 0.f;
  for (int z = 0; z < r.nsplit; ++z) acc += r.partials[((int64_t)z * r.M + m) * r.N + n];
  for (int o = 0; o < r.nouts; ++o) {
    const TnOut& t = r.outs[o];
    if (t.ptr && n >= t.col0 && n < t.col0 + t.ncols) t.ptr[(int64_t)m * t.ld + (n - t.col0)] = acc;
  }
}

int tn_nsplit(int M, int N, int64_t K) {
  int tiles = ((M + TM - 1) / TM) * ((N + TN - 1) / TN);
  int64_t want = (K + 511) / 512;            // >= 512 rows per split
  int64_t cap = (148 * 4 + tiles - 1) / tiles;  // ~4 CTAs per SM in flight
  if (cap < 1) cap = 1;
  int64_t ns = want < cap ? want : cap;
  return (int)(ns < 1 ? 1 : ns);
}

}  // namespace

size_t gemm_tn_partial_floats(int M, int N, int64_t K) { return (size_t)tn_nsplit(M, N, K) * M * N; }

int launch_gemm_nt(const RowSrc& A, int64_t B, int64_t T, int K, const float* W, int ldw, int N,
                   const float* bias, float* out, int ldo, cudaStream_t st) {
  int64_t M = B * T;
  if (M == 0) return VISDE_OK;
  RowOut o{out, T * (int64_t)ldo, (int64_t)ldo, VISDE_F32, bias};
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((N + TN - 1) / TN));
  gemm_rows_kernel<true><<<grid, NTH, 0, st>>>(A, M, T, K, W, ldw, N, o);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int launch_gemm_nn(const RowSrc& A, int64_t B, int64_t T, int K, const float* W, int ldw, int N,
                   void* out, int64_t out_bstride, int64_t out_tstride, int out_dtype,
                   cudaStream_t st) {
  int64_t M = B * T;
  if (M == 0) return VISDE_OK;
  RowOut o{out, out_bstride, out_tstride, out_dtype, nullptr};
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((N + TN - 1) / TN));
  gemm_rows_kernel<false><<<grid, NTH, 0, st>>>(A, M, T, K, W, ldw, N, o);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int launch_gemm_tn(const RowSrc& A, int M, int a_split, int a_skip, const RowSrc* Bsrc, int nsrc,
                   int64_t B, int64_t T, const TnOut* outs, int nouts, float* partials,
                   size_t partial_floats, cudaStream_t st) {
  VISDE_REQUIRE(nsrc >= 1 && nsrc <= 4 && nouts >= 1 && nouts <= 6, "gemm_tn: bad operand count");
  TnArgs g{};
  g.A = A;
  g.M = M;
  g.a_split = a_split;
  g.a_skip = a_skip;
  g.nsrc = nsrc;
  int N = 0;
  for (int s = 0; s < nsrc; ++s) {
    g.Bsrc[s] = Bsrc[s];
    N += Bsrc[s].ncols;
  }
  g.N = N;
  g.K = B * T;
  g.T = T;
  int nsplit = tn_nsplit(M, N, g.K);
  if ((size_t)nsplit * M * N > partial_floats) {
    set_error("gemm_tn: workspace too small");
    return VISDE_EWORKSPACE;
  }
  g.kchunk = (g.K + nsplit - 1) / nsplit;
  g.kchunk = (g.kchunk + TK - 1) / TK * TK;
  g.partials = partials;
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((N + TN - 1) / TN), (unsigned)nsplit);
  gemm_tn_kernel<<<grid, NTH, 0, st>>>(g);
  VISDE_CUDA_CHECK(cudaGetLastError());
  TnReduceArgs r{};
  r.partials = partials;
  r.nsplit = nsplit;
  r.M = M;
  r.N = N;
  r.nouts = nouts;
  for (int o = 0; o < nouts; ++o) r.outs[o] = outs[o];
  int total = M * N;
  gemm_tn_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(r);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}


// ------------------------------------------------------------------------------------------------
// theta pieces from sdg[b][g] = sum_t d_gi_l0[b, t, g] (both are B-row problems: two generic GEMM launches cost 60 us
// at B = 128, this one kernel ~5 us):
//   grad_theta[b][p]            = sum_g sdg[b][g] W_ih_l0[g][col0 + p]          (blocks >= G: 8 threads per b)
//   dW_ih_l0[g][col0 + p]       = sum_b sdg[b][g] theta[b][p]                   (block g: fixed-order tree over b)
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kThetaMaxP = 16;
__global__ void __launch_bounds__(256) theta_grads_kernel(const float* __restrict__ sdg, const float* __restrict__ theta,
                                                          const float* __restrict__ w_ih0, int64_t B, int G, int P, int ld0,
                                                          int col0, float* __restrict__ grad_theta, float* __restrict__ dw) {
  const int tid = threadIdx.x;
  float acc[kThetaMaxP];
#pragma unroll
  for (int q = 0; q < kThetaMaxP; ++q) acc[q] = 0.f;
  const int nrow_blocks = dw ? G : 0;
  if ((int)blockIdx.x < nrow_blocks) {
    const int g = blockIdx.x;
    for (int64_t b = tid; b < B; b += 256) {
      const float v = sdg[b * G + g];
#pragma unroll
      for (int q = 0; q < kThetaMaxP; ++q)
        if (q < P) acc[q] = fmaf(v, theta[b * P + q], acc[q]);
    }
    __shared__ float red[8][kThetaMaxP];
#pragma unroll
    for (int q = 0; q < kThetaMaxP; ++q) {
      float a = acc[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((tid & 31) == 0) red[tid >> 5][q] = a;
    }
    __syncthreads();
    if (tid < P) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += red[w][tid];
      dw[(int64_t)g * ld0 + col0 + tid] = a;
    }
  } else {
    // 32 trajectories per block; thread = (trajectory tid / 8, eighth tid % 8 of the 3H rows); W_theta staged in smem
    __shared__ float wth[768 * 2];  // [G][P] for G * P <= 1536, else read from global
    const bool staged = G * P <= 768 * 2;
    if (staged) {
      for (int idx = tid; idx < G * P; idx += 256) wth[idx] = w_ih0[(int64_t)(idx / P) * ld0 + col0 + idx % P];
    }
    __syncthreads();
    const int64_t b = (int64_t)(blockIdx.x - nrow_blocks) * 32 + (tid >> 3);
    const int part = tid & 7;
    const int g0 = part * G / 8, g1 = (part + 1) * G / 8;
    if (b < B) {
      const float* row = sdg + b * G;
#pragma unroll 8
      for (int g = g0; g < g1; ++g) {
        const float v = row[g];
        const float* w = staged ? wth + g * P : w_ih0 + (int64_t)g * ld0 + col0;
#pragma unroll
        for (int q = 0; q < kThetaMaxP; ++q)
          if (q < P) acc[q] = fmaf(v, w[q], acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < kThetaMaxP; ++q) {
      if (q < P) {  // uniform branch: P is a kernel argument
        float a = acc[q];
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        if (part == 0 && b < B) grad_theta[b * P + q] = a;
      }
    }
  }
}
}  // namespace

bool theta_grads_supported(int P) { return P >= 1 && P <= kThetaMaxP; }

// dw == nullptr: grad_theta only
int launch_theta_grads(const float* sdg, const float* theta, const float* w_ih0, int64_t B, int G, int P, int ld0, int col0,
                       float* grad_theta, float* dw, cudaStream_t st) {
  const unsigned blocks = (unsigned)((dw ? G : 0) + (B + 31) / 32);
  theta_grads_kernel<<<blocks, 256, 0, st>>>(sdg, theta, w_ih0, B, G, P, ld0, col0, grad_theta, dw);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace visde
