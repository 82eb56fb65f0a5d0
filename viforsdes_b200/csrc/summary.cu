// Posterior-sample summary: x = state_space.from_latent(z) and its mean / standard deviation over the n sampled
// trajectories at every grid point (posterior/variational_posterior.py:93-135: `result.x`, `diffusion_paths.mean(dim=0)`,
// `.std(dim=0)` with Bessel's correction), one pass over the [n, T+1, S] latent paths written by the no-stash forward.
//
// Time-parallel and HBM-bound: thread = one (t, s) column (consecutive threads = consecutive addresses, every warp
// access is a full line), grid.y = chunks of trajectories.  Each chunk keeps shifted sums (shift = its first sample,
// so the variance does not cancel), the chunks are merged in a fixed order with Chan's pairwise update: bit-identical
// run to run, no atomics.
#include "common.cuh"

namespace visde {
namespace {

constexpr int kSumThreads = 128, kSumChunkMin = 32, kSumMaxChunks = 2048;
constexpr int kSumTargetCtas = 148 * 8;  // enough CTAs to fill the machine when there are few (t, s) columns

__device__ __forceinline__ float softplus_t(float z) { return z > 20.f ? z : log1pf(expf(z)); }  // F.softplus threshold 20

struct SummaryParams {
  int64_t n, N;  // samples, columns = (T+1)*S
  int S, nchunk;
  uint32_t pos_mask;
  const float* z;
  float* x;     // [n, N] or nullptr
  float* part;  // [nchunk][3][N]: count, mean, M2
  float* mean;
  float* std;
};

__global__ void __launch_bounds__(kSumThreads) summary_partial_kernel(SummaryParams p) {
  const int64_t col = (int64_t)blockIdx.x * kSumThreads + threadIdx.x;
  if (col >= p.N) return;
  const int c = blockIdx.y;
  const int64_t per = (p.n + p.nchunk - 1) / p.nchunk;
  const int64_t i0 = c * per, i1 = i0 + per < p.n ? i0 + per : p.n;
  const bool pos = (p.pos_mask >> (int)(col % p.S)) & 1u;
  const float* __restrict__ zc = p.z + col;
  float* __restrict__ xc = p.x ? p.x + col : nullptr;
  float shift = 0.f, s1 = 0.f, s2 = 0.f;
  if (i0 < i1) {
    shift = zc[i0 * p.N];
    shift = pos ? softplus_t(shift) : shift;
  }
  // batches of 8 rows: all loads of a batch are issued before the first use (the output may not alias the input)
  constexpr int U = 8;
  for (int64_t i = i0; i < i1; i += U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = i + u < i1 ? zc[(i + u) * p.N] : 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < i1) {
        const float w = pos ? softplus_t(v[u]) : v[u];
        if (xc) xc[(i + u) * p.N] = w;
        const float d = w - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
    }
  }
  const float cnt = (float)(i1 > i0 ? i1 - i0 : 0);
  float* out = p.part + (int64_t)c * 3 * p.N;
  out[col] = cnt;
  out[p.N + col] = cnt > 0.f ? shift + s1 / cnt : 0.f;
  out[2 * p.N + col] = cnt > 0.f ? fmaxf(s2 - s1 * s1 / cnt, 0.f) : 0.f;
}

// Chan et al. pairwise update of (count, mean, M2); the order of the merges is fixed by the code below
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  if (nb == 0.f) return;
  const float nt = n + nb, delta = mb - mean;
  mean += delta * (nb / nt);
  m2 += m2b + delta * delta * (n * nb / nt);
  n = nt;
}

// one WARP per column: lane l folds chunks l, l + 32, ... in order, then a fixed shuffle tree folds the 32 lanes
// (a single thread per column walking ~600 chunks with dependent loads took 0.2 ms at n = 65 536)
__global__ void __launch_bounds__(kSumThreads) summary_merge_kernel(SummaryParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t col = (int64_t)blockIdx.x * (kSumThreads / 32) + (threadIdx.x >> 5);
  if (col >= p.N) return;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int c0 = 0; c0 < p.nchunk; c0 += 32 * 4) {
    float nb[4], mb[4], m2b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 32 * u + lane;
      const float* in = p.part + (int64_t)c * 3 * p.N;
      const bool ok = c < p.nchunk;
      nb[u] = ok ? in[col] : 0.f;
      mb[u] = ok ? in[p.N + col] : 0.f;
      m2b[u] = ok ? in[2 * p.N + col] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) chan_merge(n, mean, m2, nb[u], mb[u], m2b[u]);
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float nb = __shfl_down_sync(0xffffffffu, n, o), mb = __shfl_down_sync(0xffffffffu, mean, o);
    const float m2b = __shfl_down_sync(0xffffffffu, m2, o);
    if ((lane & (2 * o - 1)) == 0) chan_merge(n, mean, m2, nb, mb, m2b);
  }
  if (lane == 0) {
    p.mean[col] = mean;
    // torch.std(dim=0) default: unbiased; a single sample gives nan like torch
    p.std[col] = n > 1.f ? sqrtf(m2 / (n - 1.f)) : __int_as_float(0x7fc00000);
  }
}

int summary_chunks(int64_t n, int64_t N) {
  const int64_t gx = (N + kSumThreads - 1) / kSumThreads;
  int64_t cap = kSumTargetCtas / (gx > 0 ? gx : 1);
  cap = cap < 1 ? 1 : cap > kSumMaxChunks ? kSumMaxChunks : cap;
  const int64_t c = (n + kSumChunkMin - 1) / kSumChunkMin;
  return (int)(c < 1 ? 1 : c > cap ? cap : c);
}

}  // namespace
}  // namespace visde

using namespace visde;

extern "C" {

size_t visde_path_summary_workspace_bytes(int64_t n, int64_t T1, int32_t S) {
  if (n <= 0 || T1 <= 0 || S <= 0) return 256;
  return (size_t)summary_chunks(n, T1 * S) * 3 * (size_t)(T1 * S) * sizeof(float) + 256;
}

int visde_path_summary(int64_t n, int64_t T1, int32_t S, uint32_t positive_mask, const float* z, float* x,
                       float* mean, float* std, void* workspace, size_t workspace_bytes, void* stream) {
  VISDE_REQUIRE(S >= 1 && S <= VISDE_MAX_STATE, "path_summary: state dim must be in [1, %d], got %d", VISDE_MAX_STATE, S);
  VISDE_REQUIRE(n >= 0 && T1 >= 0, "path_summary: negative size");
  if (T1 == 0) return VISDE_OK;
  VISDE_REQUIRE(mean && std, "path_summary: mean / std is NULL");
  VISDE_REQUIRE(n == 0 || z, "path_summary: z is NULL");
  if (workspace_bytes < visde_path_summary_workspace_bytes(n, T1, S) || !workspace) {
    set_error("path_summary: workspace too small (%zu < %zu)", workspace_bytes, visde_path_summary_workspace_bytes(n, T1, S));
    return VISDE_EWORKSPACE;
  }
  SummaryParams p{};
  p.n = n; p.N = T1 * S; p.S = S; p.nchunk = summary_chunks(n, T1 * S); p.pos_mask = positive_mask;
  p.z = z; p.x = x; p.part = reinterpret_cast<float*>(workspace); p.mean = mean; p.std = std;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gx = (unsigned)((p.N + kSumThreads - 1) / kSumThreads);
  summary_partial_kernel<<<dim3(gx, (unsigned)p.nchunk), kSumThreads, 0, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  summary_merge_kernel<<<(unsigned)((p.N + kSumThreads / 32 - 1) / (kSumThreads / 32)), kSumThreads, 0, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // extern "C"
