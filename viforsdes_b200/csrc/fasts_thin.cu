// Thin weight-gradient reductions of the wide-state family (4 < S <= 16) in ONE time-parallel pass:
//     dW_ih_l0[:, :S] = sum_{b,t} d_gi_l0(b,t) (x) z_t(b)            [3H x S]
//     dW_out          = sum_{b,t} d_out(b,t)   (x) h_top(b,t)        [n_out x H]
//     db_out          = sum_{b,t} d_out(b,t)                         [n_out]
// (kernels/backward.py:534-590 accumulates these with global atomics inside the recurrence).  They used to be two
// generic 64 x 64-tile SIMT GEMM launches whose tiles were 84 % (N = S = 10) and 75 % (M = N = 65) empty: 3.1 of 33 ms at
// Lorenz-96 S = 10, B = 8 192.  Here every CTA streams a contiguous range of (b, t) rows through two shared-memory
// stages filled by cp.async (each row: 3H + S + n_out + H floats, read from HBM exactly once), every thread keeps its
// outputs in registers -- thread m < 3H: row m of dW_ih_l0[:, :S]; thread (g, n): rows g*RB .. of dW_out[:, n] -- and the
// per-CTA records are summed in a fixed order (deterministic, no atomics).
#include "common.cuh"

namespace visde {
namespace {

constexpr int kThinThreads = 256, kThinRows = 32, kThinMaxCtas = 592;  // 4 CTAs per SM
constexpr int kThinMaxS = 16, kThinRB = 40;                          // rows of dW_out per thread group: ceil(152 / 4) -> 40

struct ThinArgs {
  int64_t rows, T;  // rows = B * T
  int S, H, G, n_out, nop, ld0, pitch;  // nop: n_out rounded up to 4; pitch: floats per staged row
  int64_t dgrow, srow, htop_off;
  const float* dg;
  const float* paths;
  const float* dout;
  const float* stash;
  float* part;
  int64_t rows_per_cta;
  int rec;
};

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// stage `n` rows starting at global row r0 into `buf` ([kThinRows][pitch]: d_gi (G) | z (16) | d_out (nop) | h_top (H))
__device__ __forceinline__ void thin_stage(const ThinArgs& a, float* buf, int64_t r0, int n, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < n; r += kThinThreads / 32) {
    const int64_t row = r0 + r, b = row / a.T, t = row - b * a.T;
    float* dst = buf + r * a.pitch;
    const float* dg = a.dg + row * a.dgrow;
    for (int c = lane; c < a.G / 4; c += 32) cp_async16(dst + 4 * c, dg + 4 * c);
    const float* z = a.paths + (b * (a.T + 1) + t) * a.S;
    if (lane < a.S) cp_async4(dst + a.G + lane, z + lane);
    const float* dout = a.dout + row * a.n_out;
    for (int c = lane; c < a.n_out; c += 32) cp_async4(dst + a.G + kThinMaxS + c, dout + c);
    const float* h = a.stash + row * a.srow + a.htop_off;
    for (int c = lane; c < a.H / 4; c += 32) cp_async16(dst + a.G + kThinMaxS + a.nop + 4 * c, h + 4 * c);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// RBQ: float4 blocks of d_out per thread (rows of dW_out per thread group = 4 RBQ); compile time so that the hot loop has no
// run-time bounds, clamps or predicates (the first version spent 3.5x the useful instructions on them: 2.0 ms at B = 8 192)
template <int RBQ>
__global__ void __launch_bounds__(kThinThreads) fasts_thin_kernel(ThinArgs a) {
  extern __shared__ __align__(16) float thin_smem[];
  float* bufs[2] = {thin_smem, thin_smem + kThinRows * a.pitch};
  const int tid = threadIdx.x;
  const int grp = tid >> 6, n = tid & 63;
  // pad columns (z beyond S, d_out beyond n_out) are never written by the copies: zero them once
  for (int i = tid; i < 2 * kThinRows * a.pitch; i += kThinThreads) thin_smem[i] = 0.f;
  __syncthreads();
  const int64_t r_beg = (int64_t)blockIdx.x * a.rows_per_cta;
  const int64_t r_end = r_beg + a.rows_per_cta < a.rows ? r_beg + a.rows_per_cta : a.rows;
  float acc1[kThinMaxS], acc2[4 * RBQ], acc3 = 0.f;
#pragma unroll
  for (int s = 0; s < kThinMaxS; ++s) acc1[s] = 0.f;
#pragma unroll
  for (int j = 0; j < 4 * RBQ; ++j) acc2[j] = 0.f;
  constexpr int rb = 4 * RBQ;  // rows of dW_out per thread group
  const int zoff = a.G, doff = a.G + kThinMaxS, hoff = doff + a.nop;  // a.nop = 4 * rb here: the d_out region is zero-padded
  const int dmine = doff + grp * rb;

  int cur = 0;
  if (r_beg < r_end) thin_stage(a, bufs[0], r_beg, (int)(r_end - r_beg < kThinRows ? r_end - r_beg : kThinRows), tid);
  for (int64_t r0 = r_beg; r0 < r_end; r0 += kThinRows) {
    const int nrow = (int)(r_end - r0 < kThinRows ? r_end - r0 : kThinRows);
    const int64_t rn = r0 + kThinRows;
    if (rn < r_end) {
      thin_stage(a, bufs[cur ^ 1], rn, (int)(r_end - rn < kThinRows ? r_end - rn : kThinRows), tid);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* buf = bufs[cur];
    for (int k = 0; k < nrow; ++k) {
      const float* row = buf + k * a.pitch;
      if (tid < a.G) {
        const float d = row[tid];
#pragma unroll
        for (int q = 0; q < kThinMaxS / 4; ++q) {
          const float4 z = *reinterpret_cast<const float4*>(row + zoff + 4 * q);
          acc1[4 * q] = fmaf(d, z.x, acc1[4 * q]);
          acc1[4 * q + 1] = fmaf(d, z.y, acc1[4 * q + 1]);
          acc1[4 * q + 2] = fmaf(d, z.z, acc1[4 * q + 2]);
          acc1[4 * q + 3] = fmaf(d, z.w, acc1[4 * q + 3]);
        }
      }
      if (tid < a.n_out) acc3 += row[doff + tid];
      const float hv = n < a.H ? row[hoff + n] : 0.f;
#pragma unroll
      for (int q = 0; q < RBQ; ++q) {
        const float4 dv = *reinterpret_cast<const float4*>(row + dmine + 4 * q);
        acc2[4 * q] = fmaf(dv.x, hv, acc2[4 * q]);
        acc2[4 * q + 1] = fmaf(dv.y, hv, acc2[4 * q + 1]);
        acc2[4 * q + 2] = fmaf(dv.z, hv, acc2[4 * q + 2]);
        acc2[4 * q + 3] = fmaf(dv.w, hv, acc2[4 * q + 3]);
      }
    }
    __syncthreads();  // this stage is refilled two iterations from now
    cur ^= 1;
  }
  float* rec = a.part + (int64_t)blockIdx.x * a.rec;
  if (tid < a.G) {
#pragma unroll
    for (int s = 0; s < kThinMaxS; ++s)
      if (s < a.S) rec[tid * a.S + s] = acc1[s];
  }
  float* p2 = rec + a.G * a.S;
  if (n < a.H) {
#pragma unroll
    for (int j = 0; j < rb; ++j) {
      const int m = grp * rb + j;
      if (m < a.n_out) p2[m * a.H + n] = acc2[j];
    }
  }
  if (tid < a.n_out) p2[a.n_out * a.H + tid] = acc3;
}

struct ThinReduceArgs {
  const float* part;
  int ncta, rec, S, H, G, n_out, ld0;
  float* w_ih0;
  float* out_w;
  float* out_b;
};
__global__ void fasts_thin_reduce_kernel(ThinReduceArgs r) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= r.rec) return;
  float acc = 0.f;
  for (int c = 0; c < r.ncta; ++c) acc += r.part[(int64_t)c * r.rec + idx];  // fixed order
  if (idx < r.G * r.S) {
    r.w_ih0[(int64_t)(idx / r.S) * r.ld0 + idx % r.S] = acc;
  } else if (idx < r.G * r.S + r.n_out * r.H) {
    r.out_w[idx - r.G * r.S] = acc;
  } else {
    r.out_b[idx - r.G * r.S - r.n_out * r.H] = acc;
  }
}

int thin_ctas(int64_t rows) {
  int64_t c = (rows + 4 * kThinRows - 1) / (4 * kThinRows);  // at least four stages of work per CTA
  return (int)(c < 1 ? 1 : c > kThinMaxCtas ? kThinMaxCtas : c);
}
int thin_rec(int S, int H) { return 3 * H * S + (S + S * (S + 1) / 2) * (H + 1); }

}  // namespace

size_t fasts_thin_partial_floats(int64_t B, int64_t T, int S, int H) { return (size_t)thin_ctas(B * T) * thin_rec(S, H); }

int launch_fasts_thin_grads(const PathParams& p, const visde_weight_grads* gw, float* partials, size_t partial_floats,
                            cudaStream_t st) {
  const int64_t rows = p.B * p.T;
  if (rows == 0) return VISDE_OK;
  ThinArgs a{};
  const int rbq = ((p.n_out + 3) / 4 + 3) / 4;  // float4 blocks of d_out per thread group
  a.rows = rows; a.T = p.T; a.S = p.S; a.H = p.H; a.G = 3 * p.H; a.n_out = p.n_out; a.nop = 16 * rbq;
  a.ld0 = p.S + p.C + p.P;
  a.pitch = a.G + kThinMaxS + a.nop + p.H;
  a.dgrow = (int64_t)p.NL * kDgSlots * p.H; a.srow = stash_row_floats(p.NL, p.H);
  a.htop_off = ((int64_t)(p.NL - 1) * kStashSlots + kStashH) * p.H;
  a.dg = p.dg; a.paths = p.paths; a.dout = p.dout; a.stash = p.stash; a.part = partials;
  const int ncta = thin_ctas(rows);
  a.rows_per_cta = (rows + ncta - 1) / ncta;
  a.rec = thin_rec(p.S, p.H);
  VISDE_REQUIRE(p.S <= kThinMaxS && p.n_out <= 4 * kThinRB && p.H <= 64 && p.H % 4 == 0, "thin gradients: unsupported shape");
  if ((size_t)ncta * a.rec > partial_floats) {
    set_error("thin gradients: partial buffer too small (%zu < %zu)", partial_floats, (size_t)ncta * a.rec);
    return VISDE_EWORKSPACE;
  }
  void (*kern)(ThinArgs) = nullptr;
  switch (rbq) {
    case 1: case 2: kern = fasts_thin_kernel<2>; a.nop = 32; a.pitch = a.G + kThinMaxS + a.nop + p.H; break;
    case 3: kern = fasts_thin_kernel<3>; break;
    case 4: kern = fasts_thin_kernel<4>; break;
    case 5: kern = fasts_thin_kernel<5>; break;
    case 6: kern = fasts_thin_kernel<6>; break;
    case 7: kern = fasts_thin_kernel<7>; break;
    case 8: kern = fasts_thin_kernel<8>; break;
    case 9: kern = fasts_thin_kernel<9>; break;
    default: kern = fasts_thin_kernel<10>; break;
  }
  const size_t smem2 = sizeof(float) * 2 * kThinRows * a.pitch;
  static DeviceOnce attr_once[11];  // per instantiation and per device (not repeated per launch: legal under stream capture)
  int attr_dev = 0;
  DeviceOnce& once = attr_once[rbq < 2 ? 2 : rbq > 10 ? 10 : rbq];
  if (once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    once.done(attr_dev);
  }
  kern<<<ncta, kThinThreads, smem2, st>>>(a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  ThinReduceArgs r{partials, ncta, a.rec, p.S, p.H, a.G, p.n_out, a.ld0, gw->w_ih[0], gw->out_w, gw->out_b};
  fasts_thin_reduce_kernel<<<(a.rec + 255) / 256, 256, 0, st>>>(r);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace visde
