// Helpers shared by the tensor-core recurrence kernels (path_tc.cu forward, path_tc_bwd.cu backward).
#pragma once
#include <cstdlib>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace visde {
namespace {

constexpr int kTileRows = 128;
constexpr int kWTileBytes = 192 * 128;  // one weight tile: 192 gate rows x 64 fp16
constexpr int kATileBytes = 128 * 128;  // one operand tile: 128 trajectories x 64 fp16
constexpr int kOutTileBytes = 16 * 128;
constexpr int kEpiThreads = 256;  // all 8 warps of the recurrence kernels are epilogue warps
constexpr int kUPT = 32;    // hidden units per epilogue thread
constexpr int kHExp = 14;   // hidden states are scaled by 2^14 before the fp16 split

// instruction descriptor: D fp32, A/B fp16 K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_f16(int N, int M = 128) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of 16-byte chunk c (8 fp16) of row r in a K-major SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
// fp16 hi / lo halves of 8 (already scaled) floats, packed in K order
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __half2 hh = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * q] - back.x, x[2 * q + 1] - back.y);
    h[q] = *reinterpret_cast<const uint32_t*>(&hh);
    l[q] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// power-of-two exponent a with amax * 2^a in [2^13, 2^14)
__device__ __forceinline__ int scale_exp(uint32_t amax_bits) {
  if (amax_bits == 0) return 0;
  const int e = (int)(amax_bits >> 23) - 127;
  const int a = 13 - e;
  return a > 100 ? 100 : a;
}
__device__ __forceinline__ float exp2i(int a) { return __uint_as_float((uint32_t)(a + 127) << 23); }

}  // namespace
}  // namespace visde

// ---- wide-state variant of the tensor-core recurrence (4 < S <= 10, NL = 2; path_tcw.cu, path_tcw_bwd.cu) -----------------
// Shared memory cannot hold the three recurrent matrices (144 KB as fp16 hi/lo tiles) next to the 64 KB of operand tiles AND
// the S-sized tables of a wide state space, so one 48 KB buffer is time-shared between two matrices: their swizzled hi/lo
// tile images are prepared once per launch in the workspace and streamed in with cp.async.bulk (L2 -> shared memory, 96 KB
// per step and SM) behind the MMAs that read the previous occupant.
namespace visde {
constexpr int kWImg = 2 * kWTileBytes;            // hi + lo tile image of one [192 x 64] matrix: 49 152 B
constexpr int kOutRows = 80;                      // W_out tile rows: 0..63 Cholesky entries (row-major tril), 64..79 mu
constexpr int kOutImg = 2 * kOutRows * 128;       // 20 480 B
// image block in the workspace: header (ew, eo as int32) | forward images W_hh_l0, W_ih_l1, W_hh_l1 | W_out image |
// backward (transposed, K-permuted) images of the same three matrices
constexpr size_t kImgHdr = 256;
constexpr size_t kImgFwd0 = kImgHdr;
constexpr size_t kImgOut = kImgFwd0 + 3 * (size_t)kWImg;
constexpr size_t kImgBwd0 = kImgOut + kOutImg;
// backward image of W_out: B operand [64 hidden units][64 K] (fp16 hi | lo, 8 KB each); K order: Cholesky entry k for k < n_tril,
// mu component k - n_tril for the next 64 - n_tril (the remaining mu components stay on the FP32 path)
constexpr int kOutBwdImg = 2 * 64 * 128;
constexpr size_t kImgOutBwd = kImgBwd0 + 3 * (size_t)kWImg;
// backward image of the state columns of W_ih_l0: B operand [16 rows: state dim s][K = 192 in the K permutation of the
// transposed recurrent matrices], three K-blocks of [16][128 B], fp16 hi | lo; scaled by 2^ez (header word 2)
constexpr int kWzBwdImg = 2 * 3 * 16 * 128;
constexpr size_t kImgWzBwd = kImgOutBwd + kOutBwdImg;
constexpr size_t kImgBytes = kImgWzBwd + kWzBwdImg;
// (row, column) of row-major lower-triangular entry ti
__host__ __device__ constexpr int tril_row(int ti) {
  int r = 0;
  while ((r + 1) * (r + 2) / 2 <= ti) ++r;
  return r;
}
__host__ __device__ constexpr int tril_col(int ti) { return ti - tril_row(ti) * (tril_row(ti) + 1) / 2; }
// wide family: two 64-row CTAs per 128-row tile (MMA M = 64) while that still fits one CTA per SM; VISDE_TCW_M128=1 keeps M = 128
inline bool tcw_half_tiles(int64_t ntiles, int sms) {
  static const bool force128 = [] { const char* e = getenv("VISDE_TCW_M128"); return e && e[0] == '1'; }();
  return !force128 && 2 * ntiles <= sms;
}
}  // namespace visde
