// Tensor-core versions of the time-parallel GEMMs (K0, K3, K4) for sm_100a:
// tcgen05.mma kind::tf32 with FP32 accumulators in TMEM, operands staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle) straight from the caller's strided [B,T,*] tensors
// through 3-D tensor maps (out-of-range rows -- the t-1 shift, ragged T -- are zero-filled by the
// TMA unit, so there is no padding or copy pass), and a 3xTF32 split (hi*hi + hi*lo + lo*hi) done
// in shared memory by a warpgroup between the TMA and MMA stages so that results keep FP32-grade
// accuracy (|err| ~ 2^-21 relative per product; BASELINE.json asks for rtol 1e-4 in FP32).
//
//   tc_rows_kernel<N, B_MN>  one 128-row (b, t-chunk) tile per CTA, K = C or 3H:
//       K0: gi_ctx  = ctx     . Wc^T + b_ih_l0      (A K-major, B K-major,  N = 3H = 192)
//       K3: grad_ctx = d_gi_l0 . Wc                 (A K-major, B MN-major, N = C chunk of 256)
//   tc_wgrad_kernel          split-K over (b, t) with both operands MN-major (rows are k):
//       D[128, 192] += X[k, 128 cols]^T . dG[k, 192 cols]   for 5 (X, dG) pairs: ctx halves x d_gi_l0,
//       [h_l0|h_l1](t-1) x d_gh_l0, [h_l0|h_l1](t) x d_gi_l1, [h_l0|h_l1](t-1) x d_gh_l1;
//       per-CTA partials are summed in fixed order by tc_wgrad_reduce_kernel (deterministic).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = hi/lo split, then epilogue (tcgen05.ld -> global).  Two smem stages.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace visde {
namespace {

constexpr int kTcThreads = 320;  // K4 kernels: warp 0 TMA, warp 1 MMA, warps 2..9 hi/lo split (2..5 also epilogue)
constexpr int kWgtSplitThreads = 256;
constexpr uint32_t kTf32Mask = 0xffffe000u;  // keep sign, exponent and the 10 tf32 mantissa bits

// instruction descriptor: D fp32, A/B tf32, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// in place hi = trunc_tf32(x); lo = trunc_tf32(x - hi), for `nvec` float4 handled by 128 threads
__device__ __forceinline__ void split_one(const float4 x, float4& h, float4& l) {
  h.x = __uint_as_float(__float_as_uint(x.x) & kTf32Mask);
  h.y = __uint_as_float(__float_as_uint(x.y) & kTf32Mask);
  h.z = __uint_as_float(__float_as_uint(x.z) & kTf32Mask);
  h.w = __uint_as_float(__float_as_uint(x.w) & kTf32Mask);
  l.x = __uint_as_float(__float_as_uint(x.x - h.x) & kTf32Mask);
  l.y = __uint_as_float(__float_as_uint(x.y - h.y) & kTf32Mask);
  l.z = __uint_as_float(__float_as_uint(x.z - h.z) & kTf32Mask);
  l.w = __uint_as_float(__float_as_uint(x.w - h.w) & kTf32Mask);
}
// NT threads split `nvec` float4 (four independent elements in flight per thread per pass)
template <int NT = 128>
__device__ __forceinline__ void split_tile(float4* hi, float4* lo, int nvec, int tid) {
  int v = tid;
  for (; v + 3 * NT < nvec; v += 4 * NT) {
    const float4 x0 = hi[v], x1 = hi[v + NT], x2 = hi[v + 2 * NT], x3 = hi[v + 3 * NT];
    float4 h0, l0, h1, l1, h2, l2, h3, l3;
    split_one(x0, h0, l0);
    split_one(x1, h1, l1);
    split_one(x2, h2, l2);
    split_one(x3, h3, l3);
    hi[v] = h0; lo[v] = l0;
    hi[v + NT] = h1; lo[v + NT] = l1;
    hi[v + 2 * NT] = h2; lo[v + 2 * NT] = l2;
    hi[v + 3 * NT] = h3; lo[v + 3 * NT] = l3;
  }
  for (; v < nvec; v += NT) {
    float4 h, l;
    split_one(hi[v], h, l);
    hi[v] = h;
    lo[v] = l;
  }
}

// ------------------------------------------------------------------------------------------
// K0 / K3: rows kernel
// ------------------------------------------------------------------------------------------
struct RowsArgs {
  int64_t T;
  int tiles_per_b;  // ceil(T / 128)
  int num_tiles;    // B * tiles_per_b
  int num_kblocks;  // K / 32
  int n0;           // first output column handled (K3 with C > 256 launches several column chunks)
  const float* bias;
  const float* rowbias;  // tiled mode: [ceil(B/128)][N][128] added per trajectory, or nullptr
  int tiled;             // 1: a row tile is 128 TRAJECTORIES at one grid step t (tile = tb * T + t) and the output is
                         // written row-fastest, out[((tb * T + t) * N + col) * 128 + row]: the layout the tensor-core
                         // recurrence reads with one coalesced line per warp
  int a_tiled;           // 1 (K3 of the tensor-core family): A is MN-major, read from the row-fastest tiled dg
                         // [tile][t][F][128]: a row tile is 128 trajectories at one grid step (tile = tb * T + t),
                         // A slabs are [32 features][32 rows]; output rows are (b = tb * 128 + row, t)
  int a_feat_rows;       // a_tiled: features per (tile, t) block of the tiled buffer (F)
  int64_t B;             // a_tiled: trajectories (rows >= B are not stored)
  int out_tma;           // 1: fp32 output through TMA stores of swizzled [128 x 32] staging tiles (tmOut)
  void* out;
  int64_t out_bstride, out_tstride;
  int out_dtype;
  int out_cols;     // valid output columns in this chunk
};

// Persistent, warp-specialised (352 threads):
//   warp 0      A producer  : raw fp32 A tiles, NA-deep ring (the HBM stream: 16 KB per k-block)
//   warp 10     B producer  : weight hi/lo tiles (L2-resident), 2-deep ring
//   warps 2..5  split       : A raw -> hi (in place) + lo (2-deep ring shared with the B slot)
//   warp 1      MMA issuer  : 3xTF32 into one of two TMEM accumulators (tile i+1 overlaps the epilogue of i)
//   warps 6..9  epilogue    : tcgen05.ld -> (+bias) -> global
constexpr int kRowsThreads = 480;  // + warps 11..14: four more hi/lo split warps

template <int N, int NA>
struct RowsSmem {
  static constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128;
  static constexpr int OFF_ALO = NA * A_BYTES, OFF_B = OFF_ALO + 2 * A_BYTES, OFF_STG = OFF_B + 2 * 2 * B_BYTES;
  // epilogue staging tile [128 rows][32 fp32], SWIZZLE_128B: the accumulator chunk leaves through one TMA store
  // (full 128-byte row segments, rows clipped by the tensor map) instead of per-thread row-strided stores
  static constexpr int OFF_BAR = OFF_STG + A_BYTES;
  struct Bars {
    uint64_t fullA[NA], emptyA[NA], fullB[2], emptyB[2], split[2], tmemFull[2], tmemEmpty[2];
    uint32_t tmem_base;
  };
  static constexpr size_t bytes = OFF_BAR + sizeof(Bars) + 1024;
};

template <int N, bool B_MN, int NA, bool A_MN = false>
__global__ void __launch_bounds__(kRowsThreads, 1)
tc_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmOut, RowsArgs a) {
  using L = RowsSmem<N, NA>;
  constexpr int A_BYTES = L::A_BYTES, B_BYTES = L::B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * N <= 256 ? 256 : 512;
  constexpr uint32_t IDESC = make_idesc(N, A_MN, B_MN);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  typename L::Bars* bars = reinterpret_cast<typename L::Bars*>(smem + L::OFF_BAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = a.num_kblocks;
  const int ntiles = a.num_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(&bars->fullA[s], 1);
      mbar_init(&bars->emptyA[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->fullB[s], 1);
      mbar_init(&bars->emptyB[s], 1);
      mbar_init(&bars->split[s], 256);
      mbar_init(&bars->tmemFull[s], 1);
      mbar_init(&bars->tmemEmpty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_b, t0 = (tile % a.tiles_per_b) * 128;
        for (int kb = 0; kb < nk; ++kb, ++g) {
          const uint32_t sa = g % NA;
          if (g >= NA) mbar_wait(&bars->emptyA[sa], ((g / NA) - 1) & 1);
          mbar_expect_tx(&bars->fullA[sa], A_BYTES);
          if (A_MN) {
#pragma unroll
            for (int sl = 0; sl < 4; ++sl)
              tma_load_2d(smem + sa * A_BYTES + sl * 4096, &tmA, &bars->fullA[sa], sl * 32, tile * a.a_feat_rows + kb * 32);
          } else if (a.tiled)
            tma_load_3d(smem + sa * A_BYTES, &tmA, &bars->fullA[sa], kb * 32, (int)(tile % a.T), (int)(tile / a.T) * 128);
          else
            tma_load_3d(smem + sa * A_BYTES, &tmA, &bars->fullA[sa], kb * 32, t0, b);
        }
      }
    }
  } else if (warp == 10) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kb = 0; kb < nk; ++kb, ++g) {
          const uint32_t sb = g & 1;
          if (g >= 2) mbar_wait(&bars->emptyB[sb], ((g >> 1) - 1) & 1);
          uint8_t* bh = smem + L::OFF_B + sb * 2 * B_BYTES;
          mbar_expect_tx(&bars->fullB[sb], 2 * B_BYTES);
          if (B_MN) {
#pragma unroll
            for (int sl = 0; sl < N / 32; ++sl) {
              tma_load_2d(bh + sl * 4096, &tmBhi, &bars->fullB[sb], a.n0 + sl * 32, kb * 32);
              tma_load_2d(bh + B_BYTES + sl * 4096, &tmBlo, &bars->fullB[sb], a.n0 + sl * 32, kb * 32);
            }
          } else {
            tma_load_2d(bh, &tmBhi, &bars->fullB[sb], kb * 32, a.n0);
            tma_load_2d(bh + B_BYTES, &tmBlo, &bars->fullB[sb], kb * 32, a.n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t g = 0, lt = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
        const uint32_t acc = lt & 1;
        if (lt >= 2) mbar_wait(&bars->tmemEmpty[acc], ((lt >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t dcol = tmem_d + acc * N;
        for (int kb = 0; kb < nk; ++kb, ++g) {
          const uint32_t sa = g % NA, sb = g & 1, ph = (g >> 1) & 1;
          mbar_wait(&bars->fullB[sb], ph);
          mbar_wait(&bars->split[sb], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + sa * A_BYTES), a_lo = smem_u32(smem + L::OFF_ALO + sb * A_BYTES);
          const uint32_t b_hi = smem_u32(smem + L::OFF_B + sb * 2 * B_BYTES), b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t dah = A_MN ? desc_mnmajor(a_hi, j) : desc_kmajor(a_hi, j);
            const uint64_t dal = A_MN ? desc_mnmajor(a_lo, j) : desc_kmajor(a_lo, j);
            const uint64_t dbh = B_MN ? desc_mnmajor(b_hi, j) : desc_kmajor(b_hi, j);
            const uint64_t dbl = B_MN ? desc_mnmajor(b_lo, j) : desc_kmajor(b_lo, j);
            umma_tf32(dcol, dal, dbh, IDESC, (kb | j) != 0);  // small terms first
            umma_tf32(dcol, dah, dbl, IDESC, 1);
            umma_tf32(dcol, dah, dbh, IDESC, 1);
          }
          umma_commit(&bars->emptyA[sa]);
          umma_commit(&bars->emptyB[sb]);
        }
        umma_commit(&bars->tmemFull[acc]);
      }
    }
  } else if ((warp >= 2 && warp <= 5) || warp >= 11) {
    const int tid128 = warp <= 5 ? threadIdx.x - 64 : threadIdx.x - 352 + 128;  // 0..255
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int kb = 0; kb < nk; ++kb, ++g) {
        const uint32_t sa = g % NA, sb = g & 1;
        mbar_wait(&bars->fullA[sa], (g / NA) & 1);
        if (g >= 2) mbar_wait(&bars->emptyB[sb], ((g >> 1) - 1) & 1);  // lo slot released by the MMAs of g-2
        split_tile<256>(reinterpret_cast<float4*>(smem + sa * A_BYTES),
                        reinterpret_cast<float4*>(smem + L::OFF_ALO + sb * A_BYTES), A_BYTES / 16, tid128);
        fence_proxy_async();
        mbar_arrive(&bars->split[sb]);
      }
    }
  } else if (warp >= 6 && warp <= 9) {
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int b = tile / a.tiles_per_b, t0 = (tile % a.tiles_per_b) * 128;
      const uint32_t acc = lt & 1;
      mbar_wait(&bars->tmemFull[acc], (lt >> 1) & 1);
      tc_fence_after();
      int t = t0 + row;
      bool row_ok = t < a.T;
      int64_t obase = (int64_t)b * a.out_bstride + (int64_t)t * a.out_tstride + a.n0;
      if (A_MN) {
        const int64_t bb = (int64_t)(tile / a.T) * 128 + row;
        t = (int)(tile % a.T);
        row_ok = bb < a.B;
        obase = bb * a.out_bstride + (int64_t)t * a.out_tstride + a.n0;
      }
      if (a.tiled) {
        // rows are trajectories tb * 128 + row at grid step t; fp32 row-fastest output, always in bounds (padded)
        const int64_t tb = tile / a.T;
        float* o = reinterpret_cast<float*>(a.out) + (int64_t)tile * N * 128 + row;
        const float* rbt = a.rowbias ? a.rowbias + tb * N * 128 + row : nullptr;
#pragma unroll 1
        for (int c = 0; c < N / 32; ++c) {
          float v[32];
          tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + acc * N + c * 32, v);
          if (rbt) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += rbt[(c * 32 + q) * 128];
          }
#pragma unroll
          for (int q = 0; q < 32; ++q) o[(c * 32 + q) * 128] = v[q];
        }
        tc_fence_before();
        mbar_arrive(&bars->tmemEmpty[acc]);
        continue;
      }
      if (a.out_tma) {
        uint8_t* stg = smem + L::OFF_STG;
#pragma unroll 1
        for (int c = 0; c < a.out_cols / 32; ++c) {
          float v[32];
          tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + acc * N + c * 32, v);
          if (a.bias) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += a.bias[a.n0 + c * 32 + q];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)  // 16-byte chunk j of row `row` lives at chunk j ^ (row % 8): conflict-free
            *reinterpret_cast<float4*>(stg + row * 128 + ((j ^ (row & 7)) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (warp == 6 && lane == 0) {
            if (A_MN) tma_store_3d(&tmOut, stg, a.n0 + c * 32, (int)(tile % a.T), (int)(tile / a.T) * 128);
            else tma_store_3d(&tmOut, stg, a.n0 + c * 32, t0, b);
            tma_store_commit_and_wait_read();
          }
          named_bar_sync(1, 128);  // staging tile reusable
        }
        tc_fence_before();
        mbar_arrive(&bars->tmemEmpty[acc]);
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < N / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + acc * N + c * 32, v);
        if (a.bias) {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] += a.bias[a.n0 + c * 32 + q];
        }
        if (row_ok) {
          if (a.out_dtype == VISDE_BF16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + obase + c * 32;
            if (c * 32 + 32 <= a.out_cols && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
              for (int q = 0; q < 32; q += 8) {
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[q], v[q + 1]), p1 = __floats2bfloat162_rn(v[q + 2], v[q + 3]);
                const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[q + 4], v[q + 5]), p3 = __floats2bfloat162_rn(v[q + 6], v[q + 7]);
                *reinterpret_cast<uint4*>(o + q) =
                    make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                               *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
              }
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (c * 32 + q < a.out_cols) o[q] = __float2bfloat16(v[q]);
            }
          } else {
            float* o = reinterpret_cast<float*>(a.out) + obase + c * 32;
            if (c * 32 + 32 <= a.out_cols && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
              for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(o + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (c * 32 + q < a.out_cols) o[q] = v[q];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->tmemEmpty[acc]);
    }
    if (a.out_tma && warp == 6 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// K4: split-K weight-gradient kernel (both operands MN-major)
// ------------------------------------------------------------------------------------------
struct WgProblem {
  int a_map, b_map;  // which tensor map (0 = ctx, 1 = dg, 2 = stash)
  int a_cols[4];     // column coordinate of each 32-wide A slab (M = 128)
  int b_cols[6];     // column coordinate of each 32-wide B slab (N = 192)
  int a_tshift, b_tshift;
};
struct WgArgs {
  WgProblem prob[5];
  int nprob, nsplit;
  int64_t total_kblocks;  // B * ceil(T / 32)
  int kb_per_b;           // ceil(T / 32)
  float* partials;        // [nprob][nsplit][128][192]
};

// Pipeline (both K4 kernels): raw fp32 tiles land in a 3-deep ring (TMA runs two k-blocks ahead of the MMAs); the split
// warps turn a raw stage into its hi part in place and write the lo part into a 2-deep ring; the MMA warp frees both.
constexpr int kWgtRaw = 3, kWgtLo = 2;
struct WgtBarriers {
  uint64_t full[kWgtRaw], emptyRaw[kWgtRaw], split[kWgtLo], emptyLo[kWgtLo], accum;
  uint32_t tmem_base;
};
constexpr int kWgtStageBytes = 128 * 128 + 192 * 128;  // A [128 x 32 fp32] + B [192 x 32 fp32] = 40 KB
constexpr size_t kWgtSmemBytes = (size_t)(kWgtRaw + kWgtLo) * kWgtStageBytes + sizeof(WgtBarriers) + 1024;

__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                const __grid_constant__ CUtensorMap tm2, WgArgs a) {
  constexpr int N = 192;
  constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128;
  constexpr uint32_t TMEM_COLS = 256;
  constexpr uint32_t IDESC = make_idesc(N, true, true);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* lo_ring = smem + kWgtRaw * kWgtStageBytes;
  WgtBarriers* bars = reinterpret_cast<WgtBarriers*>(smem + (kWgtRaw + kWgtLo) * kWgtStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pi = blockIdx.y, split = blockIdx.x;
  const WgProblem& pr = a.prob[pi];
  const int64_t per = (a.total_kblocks + a.nsplit - 1) / a.nsplit;
  const int64_t kb0 = split * per;
  const int64_t kb1 = kb0 + per < a.total_kblocks ? kb0 + per : a.total_kblocks;
  const int nk = kb1 > kb0 ? (int)(kb1 - kb0) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgtRaw; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->emptyRaw[s], 1);
    }
    for (int s = 0; s < kWgtLo; ++s) {
      mbar_init(&bars->split[s], kWgtSplitThreads);
      mbar_init(&bars->emptyLo[s], 1);
    }
    mbar_init(&bars->accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mA = pr.a_map == 0 ? &tm0 : pr.a_map == 1 ? &tm1 : &tm2;
      const CUtensorMap* mB = pr.b_map == 0 ? &tm0 : pr.b_map == 1 ? &tm1 : &tm2;
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kWgtRaw;
        if (kb >= kWgtRaw) mbar_wait(&bars->emptyRaw[s], ((kb / kWgtRaw) - 1) & 1);
        const int64_t g = kb0 + kb;
        const int b = (int)(g / a.kb_per_b), t32 = (int)(g % a.kb_per_b) * 32;
        uint8_t* st = smem + s * kWgtStageBytes;
        mbar_expect_tx(&bars->full[s], A_BYTES + B_BYTES);
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)
          tma_load_3d(st + sl * 4096, mA, &bars->full[s], pr.a_cols[sl], t32 + pr.a_tshift, b);
#pragma unroll
        for (int sl = 0; sl < 6; ++sl)
          tma_load_3d(st + A_BYTES + sl * 4096, mB, &bars->full[s], pr.b_cols[sl], t32 + pr.b_tshift, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kWgtRaw, sl = kb % kWgtLo;
        mbar_wait(&bars->split[sl], (kb / kWgtLo) & 1);
        tc_fence_after();
        const uint32_t hi = smem_u32(smem + s * kWgtStageBytes), lo = smem_u32(lo_ring + sl * kWgtStageBytes);
        const uint32_t a_hi = hi, a_lo = lo, b_hi = hi + A_BYTES, b_lo = lo + A_BYTES;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t dah = desc_mnmajor(a_hi, j), dal = desc_mnmajor(a_lo, j);
          const uint64_t dbh = desc_mnmajor(b_hi, j), dbl = desc_mnmajor(b_lo, j);
          umma_tf32(tmem_d, dal, dbh, IDESC, (kb | j) != 0);
          umma_tf32(tmem_d, dah, dbl, IDESC, 1);
          umma_tf32(tmem_d, dah, dbh, IDESC, 1);
        }
        umma_commit(&bars->emptyRaw[s]);
        umma_commit(&bars->emptyLo[sl]);
      }
      if (nk > 0) umma_commit(&bars->accum); else mbar_arrive(&bars->accum);
    }
  } else {
    const int tid128 = threadIdx.x - 64;
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % kWgtRaw, sl = kb % kWgtLo;
      mbar_wait(&bars->full[s], (kb / kWgtRaw) & 1);
      if (kb >= kWgtLo) mbar_wait(&bars->emptyLo[sl], ((kb / kWgtLo) - 1) & 1);
      split_tile<kWgtSplitThreads>(reinterpret_cast<float4*>(smem + s * kWgtStageBytes),
                                   reinterpret_cast<float4*>(lo_ring + sl * kWgtStageBytes), kWgtStageBytes / 16, tid128);
      fence_proxy_async();
      mbar_arrive(&bars->split[sl]);
    }
    if (warp <= 5) {
    mbar_wait(&bars->accum, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    float* out = a.partials + (((int64_t)pi * a.nsplit + split) * 128 + row) * N;
#pragma unroll 1
    for (int c = 0; c < N / 32; ++c) {
      float v[32];
      if (nk > 0) {
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + c * 32, v);
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 32; q += 4)
        *reinterpret_cast<float4*>(out + c * 32 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}

// K4 of the tensor-core family: the same split-K products read from the row-fastest tiled buffers
// dg [tile][t][F][128] / stash [tile][t][Fs][128] (K = trajectories of a tile: both operands K-major, 4-D tensor
// maps so that the t-1 shift of h(t-1) is an out-of-bounds zero fill) and from the caller's ctx [B,T,C]
// (MN-major slabs, K = 32 trajectories at one grid step).  One k-block = 32 trajectories of one (tile, t).
struct WgtProblem {
  int a_ctx;       // 1: A = 128 ctx columns (MN-major), 0: A = [h_l0 | h_l1] from the tiled stash (K-major)
  int a_cols[4];   // ctx: column of each 32-wide slab; stash: feature row of h_l0, h_l1
  int a_tshift;
  int b_feat[3];   // feature rows of the three 64-wide gate groups in dg
};
struct WgtArgs {
  WgtProblem prob[5];
  int nprob, nsplit;
  int64_t total_kblocks;  // ntile * T * 4
  int T;
  float* partials;        // [nprob][nsplit][128][192]
};

__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad_tiled_kernel(const __grid_constant__ CUtensorMap tmCtx, const __grid_constant__ CUtensorMap tmDg,
                      const __grid_constant__ CUtensorMap tmSt, WgtArgs a) {
  constexpr int N = 192;
  constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128;
  constexpr uint32_t TMEM_COLS = 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* lo_ring = smem + kWgtRaw * kWgtStageBytes;
  WgtBarriers* bars = reinterpret_cast<WgtBarriers*>(smem + (kWgtRaw + kWgtLo) * kWgtStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pi = blockIdx.y, split = blockIdx.x;
  const WgtProblem& pr = a.prob[pi];
  const int64_t per = (a.total_kblocks + a.nsplit - 1) / a.nsplit;
  const int64_t kb0 = split * per;
  const int64_t kb1 = kb0 + per < a.total_kblocks ? kb0 + per : a.total_kblocks;
  const int nk = kb1 > kb0 ? (int)(kb1 - kb0) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgtRaw; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->emptyRaw[s], 1);
    }
    for (int s = 0; s < kWgtLo; ++s) {
      mbar_init(&bars->split[s], kWgtSplitThreads);
      mbar_init(&bars->emptyLo[s], 1);
    }
    mbar_init(&bars->accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kWgtRaw;
        if (kb >= kWgtRaw) mbar_wait(&bars->emptyRaw[s], ((kb / kWgtRaw) - 1) & 1);
        const int64_t g = kb0 + kb;
        const int kq = (int)(g & 3), t = (int)((g >> 2) % a.T), tb = (int)((g >> 2) / a.T);
        uint8_t* st = smem + s * kWgtStageBytes;
        mbar_expect_tx(&bars->full[s], A_BYTES + B_BYTES);
        if (pr.a_ctx) {
#pragma unroll
          for (int sl = 0; sl < 4; ++sl)
            tma_load_3d(st + sl * 4096, &tmCtx, &bars->full[s], pr.a_cols[sl], t, tb * 128 + kq * 32);
        } else {
          tma_load_4d(st, &tmSt, &bars->full[s], kq * 32, pr.a_cols[0], t + pr.a_tshift, tb);
          tma_load_4d(st + 8192, &tmSt, &bars->full[s], kq * 32, pr.a_cols[1], t + pr.a_tshift, tb);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
          tma_load_4d(st + A_BYTES + q * 8192, &tmDg, &bars->full[s], kq * 32, pr.b_feat[q], t, tb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(N, pr.a_ctx != 0, false);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % kWgtRaw, sl = kb % kWgtLo;
        mbar_wait(&bars->split[sl], (kb / kWgtLo) & 1);
        tc_fence_after();
        const uint32_t hi = smem_u32(smem + s * kWgtStageBytes), lo = smem_u32(lo_ring + sl * kWgtStageBytes);
        const uint32_t a_hi = hi, a_lo = lo, b_hi = hi + A_BYTES, b_lo = lo + A_BYTES;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t dah = pr.a_ctx ? desc_mnmajor(a_hi, j) : desc_kmajor(a_hi, j);
          const uint64_t dal = pr.a_ctx ? desc_mnmajor(a_lo, j) : desc_kmajor(a_lo, j);
          const uint64_t dbh = desc_kmajor(b_hi, j), dbl = desc_kmajor(b_lo, j);
          umma_tf32(tmem_d, dal, dbh, idesc, (kb | j) != 0);
          umma_tf32(tmem_d, dah, dbl, idesc, 1);
          umma_tf32(tmem_d, dah, dbh, idesc, 1);
        }
        umma_commit(&bars->emptyRaw[s]);
        umma_commit(&bars->emptyLo[sl]);
      }
      if (nk > 0) umma_commit(&bars->accum); else mbar_arrive(&bars->accum);
    }
  } else {
    const int tid128 = threadIdx.x - 64;
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % kWgtRaw, sl = kb % kWgtLo;
      mbar_wait(&bars->full[s], (kb / kWgtRaw) & 1);
      if (kb >= kWgtLo) mbar_wait(&bars->emptyLo[sl], ((kb / kWgtLo) - 1) & 1);
      split_tile<kWgtSplitThreads>(reinterpret_cast<float4*>(smem + s * kWgtStageBytes),
                                   reinterpret_cast<float4*>(lo_ring + sl * kWgtStageBytes), kWgtStageBytes / 16, tid128);
      fence_proxy_async();
      mbar_arrive(&bars->split[sl]);
    }
    if (warp <= 5) {
    mbar_wait(&bars->accum, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    float* out = a.partials + (((int64_t)pi * a.nsplit + split) * 128 + row) * N;
#pragma unroll 1
    for (int c = 0; c < N / 32; ++c) {
      float v[32];
      if (nk > 0) {
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + c * 32, v);
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 32; q += 4)
        *reinterpret_cast<float4*>(out + c * 32 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}

// fixed-order sum over splits + scatter (transposed) into the gradient tensors
struct WgScatter {
  float* dst;      // dst[n * ld + col0 + (row - row0)] = sum_split partial[row][n]
  int ld, col0, row0, nrows;
};
struct WgReduceArgs {
  const float* partials;
  int nsplit;
  WgScatter sc[5];
  int nprob;
};
__global__ void tc_wgrad_reduce_kernel(WgReduceArgs r) {
  const int pi = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over 128 * 192, n fastest
  if (idx >= 128 * 192) return;
  const int row = idx / 192, n = idx % 192;
  const WgScatter& s = r.sc[pi];
  if (row < s.row0 || row >= s.row0 + s.nrows) return;
  const float* p = r.partials + (int64_t)pi * r.nsplit * 128 * 192 + idx;
  float acc = 0.f;
  for (int z = 0; z < r.nsplit; ++z) acc += p[(int64_t)z * 128 * 192];
  s.dst[(int64_t)n * s.ld + s.col0 + (row - s.row0)] = acc;
}

// hi/lo split of the context columns of W_ih_l0, packed [3H, C] (K0: K-major B; K3: MN-major B)
__global__ void split_weights_kernel(const float* __restrict__ w, int ld, int col0, int rows, int cols, float* hi,
                                     float* lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx % cols;
  const float x = w[(int64_t)r * ld + col0 + c];
  const float h = __uint_as_float(__float_as_uint(x) & kTf32Mask);
  hi[idx] = h;
  lo[idx] = __uint_as_float(__float_as_uint(x - h) & kTf32Mask);
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp32 tensor [d2][d1][d0] (d0 contiguous), strides in elements; box [b1 rows of dim1][b0 cols of dim0]
int make_map(CUtensorMap* m, const void* base, int rank, const int64_t* dims, const int64_t* strides_elems,
             const int* box, bool mn_major = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return VISDE_ECUDA;
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = (cuuint64_t)strides_elems[i] * sizeof(float);
  CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return VISDE_ECUDA;
  }
  return VISDE_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int N, bool B_MN, int NA, bool A_MN = false>
int launch_rows(const CUtensorMap& mA, const CUtensorMap& mBh, const CUtensorMap& mBl, const CUtensorMap& mOut, RowsArgs a,
                int64_t B, cudaStream_t st) {
  const size_t smem = RowsSmem<N, NA>::bytes;
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(tc_rows_kernel<N, B_MN, NA, A_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  a.num_tiles = (a.tiled || a.a_tiled) ? (int)(((B + 127) / 128) * a.T) : (int)(B * a.tiles_per_b);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.num_tiles < sms ? a.num_tiles : sms;
  tc_rows_kernel<N, B_MN, NA, A_MN><<<grid, kRowsThreads, smem, st>>>(mA, mBh, mBl, mOut, a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// public (internal) entry points
// ------------------------------------------------------------------------------------------
bool tc_supported(int H, int NL, int C, const visde_ctx_view* ctx) {
  if (H != 64 || NL > 2 || (C != 128 && C != 256)) return false;
  if (!ctx || ctx->dtype != VISDE_F32 || !aligned16(ctx->ptr)) return false;
  if (ctx->batch_stride % 4 != 0 || ctx->time_stride % 4 != 0) return false;
  return get_encode() != nullptr;
}

size_t tc_weight_scratch_floats(int H, int C) { return (size_t)2 * 3 * H * C; }
size_t tc_wgrad_partial_floats(int NL, int C) {
  const int nprob = C / 128 + (NL == 1 ? 1 : 3);
  const int nsplit = 148 / nprob;
  return (size_t)nprob * nsplit * 128 * 192;
}

// packed hi / lo copies of W_ih_l0[:, S:S+C] into scratch (hi at scratch, lo at scratch + 3H*C)
int tc_split_weights(const float* w_ih0, int ld0, int S, int H, int C, float* scratch, cudaStream_t st) {
  const int n = 3 * H * C;
  split_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w_ih0, ld0, S, 3 * H, C, scratch, scratch + n);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

// K0: gi_ctx[B,T,192] = ctx . Wc^T + b_ih0
int tc_ctx_proj(const visde_ctx_view* ctx, int64_t B, int64_t T, int C, int H, const float* wsplit, const float* bias,
                const float* rowbias_tiled, float* gi_ctx, bool tiled, cudaStream_t st) {
  CUtensorMap mA, mBh, mBl;
  const int64_t dA[3] = {C, T, B}, sA[2] = {ctx->time_stride, ctx->batch_stride};
  const int boxA[3] = {32, tiled ? 1 : 128, tiled ? 128 : 1};
  int rc = make_map(&mA, ctx->ptr, 3, dA, sA, boxA);
  if (rc) return rc;
  const int64_t dB[2] = {C, 3 * H}, sB[1] = {C};
  const int boxB[2] = {32, 3 * H};
  if ((rc = make_map(&mBh, wsplit, 2, dB, sB, boxB))) return rc;
  if ((rc = make_map(&mBl, wsplit + (size_t)3 * H * C, 2, dB, sB, boxB))) return rc;
  RowsArgs a{};
  a.T = T;
  a.tiles_per_b = (int)((T + 127) / 128);
  a.num_kblocks = C / 32;
  a.n0 = 0;
  a.bias = bias;
  a.rowbias = rowbias_tiled;
  a.tiled = tiled ? 1 : 0;
  a.out = gi_ctx;
  a.out_bstride = T * (int64_t)(3 * H);
  a.out_tstride = 3 * H;
  a.out_dtype = VISDE_F32;
  a.out_cols = 3 * H;
  CUtensorMap mOut = mA;  // placeholder when the TMA-store epilogue is off
  if (!tiled) {
    const int64_t dO[3] = {3 * H, T, B}, sO[2] = {3 * H, T * (int64_t)(3 * H)};
    const int boxO[3] = {32, 128, 1};
    if ((rc = make_map(&mOut, gi_ctx, 3, dO, sO, boxO))) return rc;
    a.out_tma = 1;
  }
  return launch_rows<192, false, 4>(mA, mBh, mBl, mOut, a, B, st);
}

// K3: grad_ctx[b,t,:] = d_gi_l0[(b,t), :192] . Wc   (dg rows have `dg_row` floats).  dg_tiled: dg is the row-fastest
// tiled buffer [ceil(B/128)][T][dg_row][128] of the tensor-core backward (A operand MN-major)
int tc_grad_ctx(const float* dg, int64_t dg_row, int64_t B, int64_t T, int C, int H, const float* wsplit,
                const visde_ctx_grad_view* out, bool dg_tiled, cudaStream_t st) {
  CUtensorMap mA, mBh, mBl;
  int rc;
  if (dg_tiled) {
    const int64_t dA[2] = {128, ((B + 127) / 128) * T * dg_row}, sA[1] = {128};
    const int boxA[2] = {32, 32};
    rc = make_map(&mA, dg, 2, dA, sA, boxA, true);
  } else {
    const int64_t dA[3] = {3 * H, T, B}, sA[2] = {dg_row, T * dg_row};
    const int boxA[3] = {32, 128, 1};
    rc = make_map(&mA, dg, 3, dA, sA, boxA);
  }
  if (rc) return rc;
  const int64_t dB[2] = {C, 3 * H}, sB[1] = {C};
  const int boxB[2] = {32, 32};
  if ((rc = make_map(&mBh, wsplit, 2, dB, sB, boxB, true))) return rc;
  if ((rc = make_map(&mBl, wsplit + (size_t)3 * H * C, 2, dB, sB, boxB, true))) return rc;
  // fp32 grad_ctx with 16-byte aligned strides leaves through TMA stores (rows beyond T / B are clipped by the map)
  CUtensorMap mOut = mA;
  const bool out_tma = out->dtype == VISDE_F32 && aligned16(out->ptr) && out->batch_stride % 4 == 0 &&
                       out->time_stride % 4 == 0 && C % 32 == 0;
  if (out_tma) {
    const int64_t dO[3] = {C, T, B}, sO[2] = {out->time_stride, out->batch_stride};
    const int boxT[3] = {32, 1, 128}, boxS[3] = {32, 128, 1};
    if ((rc = make_map(&mOut, out->ptr, 3, dO, sO, dg_tiled ? boxT : boxS))) return rc;
  }
  for (int n0 = 0; n0 < C; n0 += 256) {
    RowsArgs a{};
    a.T = T;
    a.out_tma = out_tma ? 1 : 0;
    a.tiles_per_b = (int)((T + 127) / 128);
    a.num_kblocks = 3 * H / 32;
    a.n0 = n0;
    a.bias = nullptr;
    a.a_tiled = dg_tiled ? 1 : 0;
    a.a_feat_rows = (int)dg_row;
    a.B = B;
    a.out = out->ptr;
    a.out_bstride = out->batch_stride;
    a.out_tstride = out->time_stride;
    a.out_dtype = out->dtype;
    a.out_cols = C - n0 < 256 ? C - n0 : 256;
    if (dg_tiled)
      rc = a.out_cols == 256 ? launch_rows<256, true, 3, true>(mA, mBh, mBl, mOut, a, B, st)
                             : launch_rows<128, true, 4, true>(mA, mBh, mBl, mOut, a, B, st);
    else
      rc = a.out_cols == 256 ? launch_rows<256, true, 3>(mA, mBh, mBl, mOut, a, B, st)
                             : launch_rows<128, true, 4>(mA, mBh, mBl, mOut, a, B, st);
    if (rc) return rc;
  }
  return VISDE_OK;
}

// K4 (big part): dW_ih0[:, S:S+C], dW_hh0, dW_ih1, dW_hh1 from ctx, dg and the h slots of the stash
int tc_wgrads(const visde_ctx_view* ctx, const float* dg, const float* stash, int64_t B, int64_t T, int S, int C, int P,
              int H, int NL, const visde_weight_grads* gw, float* partials, size_t partial_floats, cudaStream_t st) {
  const int64_t dg_row = (int64_t)NL * kDgSlots * H, st_row = stash_row_floats(NL, H);
  CUtensorMap m0, m1, m2;
  const int box[3] = {32, 32, 1};
  {
    const int64_t d[3] = {C, T, B}, s[2] = {ctx->time_stride, ctx->batch_stride};
    int rc = make_map(&m0, ctx->ptr, 3, d, s, box, true);
    if (rc) return rc;
  }
  {
    const int64_t d[3] = {dg_row, T, B}, s[2] = {dg_row, T * dg_row};
    int rc = make_map(&m1, dg, 3, d, s, box, true);
    if (rc) return rc;
  }
  {
    const int64_t d[3] = {st_row, T, B}, s[2] = {st_row, T * st_row};
    int rc = make_map(&m2, stash, 3, d, s, box, true);
    if (rc) return rc;
  }
  WgArgs a{};
  WgReduceArgs r{};
  int np = 0;
  const int ld0 = S + C + P;
  auto set_b = [&](WgProblem& p, int layer, bool gh) {
    const int base = layer * kDgSlots * H;
    for (int sl = 0; sl < 6; ++sl) {
      int col = sl * 32;
      if (gh && col >= 2 * H) col += H;  // (r, u, n_hh) slots
      p.b_cols[sl] = base + col;
    }
    p.b_map = 1;
    p.b_tshift = 0;
  };
  auto hcat = [&](WgProblem& p, int tshift) {
    const int h0 = kStashH * H, h1 = (NL > 1 ? kStashSlots + kStashH : kStashH) * H;
    p.a_map = 2;
    p.a_cols[0] = h0;
    p.a_cols[1] = h0 + 32;
    p.a_cols[2] = h1;
    p.a_cols[3] = h1 + 32;
    p.a_tshift = tshift;
  };
  for (int c0 = 0; c0 < C; c0 += 128) {  // dW_ih0[:, S + c0 : S + c0 + 128]
    WgProblem& p = a.prob[np];
    p.a_map = 0;
    for (int sl = 0; sl < 4; ++sl) p.a_cols[sl] = c0 + sl * 32;
    p.a_tshift = 0;
    set_b(p, 0, false);
    r.sc[np] = WgScatter{gw->w_ih[0], ld0, S + c0, 0, 128};
    ++np;
  }
  {  // dW_hh0 = d_gh0^T h0(t-1)
    WgProblem& p = a.prob[np];
    hcat(p, -1);
    set_b(p, 0, true);
    r.sc[np] = WgScatter{gw->w_hh[0], H, 0, 0, H};
    ++np;
  }
  if (NL > 1) {
    {  // dW_ih1 = d_gi1^T h0(t)
      WgProblem& p = a.prob[np];
      hcat(p, 0);
      set_b(p, 1, false);
      r.sc[np] = WgScatter{gw->w_ih[1], H, 0, 0, H};
      ++np;
    }
    {  // dW_hh1 = d_gh1^T h1(t-1)
      WgProblem& p = a.prob[np];
      hcat(p, -1);
      set_b(p, 1, true);
      r.sc[np] = WgScatter{gw->w_hh[1], H, 0, H, H};
      ++np;
    }
  }
  a.nprob = np;
  a.nsplit = 148 / np;
  a.kb_per_b = (int)((T + 31) / 32);
  a.total_kblocks = B * a.kb_per_b;
  if (a.total_kblocks < a.nsplit) a.nsplit = (int)a.total_kblocks;
  a.partials = partials;
  if ((size_t)np * a.nsplit * 128 * 192 > partial_floats) {
    set_error("tc_wgrads: workspace too small");
    return VISDE_EWORKSPACE;
  }
  const size_t smem = kWgtSmemBytes;
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  tc_wgrad_kernel<<<dim3(a.nsplit, np), kTcThreads, smem, st>>>(m0, m1, m2, a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  r.partials = partials;
  r.nsplit = a.nsplit;
  r.nprob = np;
  tc_wgrad_reduce_kernel<<<dim3((128 * 192 + 255) / 256, np), 256, 0, st>>>(r);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

// K4 (big part) of the tensor-core family: dg / stash are the row-fastest tiled buffers
int tc_wgrads_tiled(const visde_ctx_view* ctx, const float* dg, const float* stash, int64_t B, int64_t T, int S, int C,
                    int P, int H, int NL, const visde_weight_grads* gw, float* partials, size_t partial_floats,
                    cudaStream_t st) {
  const int64_t F = (int64_t)NL * kDgSlots * H, Fs = stash_row_floats(NL, H), ntile = (B + 127) / 128;
  CUtensorMap m0, m1, m2;
  {
    const int64_t d[3] = {C, T, B}, s[2] = {ctx->time_stride, ctx->batch_stride};
    const int box[3] = {32, 1, 32};
    int rc = make_map(&m0, ctx->ptr, 3, d, s, box, true);
    if (rc) return rc;
  }
  {
    const int64_t d[4] = {128, F, T, ntile}, s[3] = {128, F * 128, T * F * 128};
    const int box[4] = {32, 64, 1, 1};
    int rc = make_map(&m1, dg, 4, d, s, box);
    if (rc) return rc;
  }
  {
    const int64_t d[4] = {128, Fs, T, ntile}, s[3] = {128, Fs * 128, T * Fs * 128};
    const int box[4] = {32, 64, 1, 1};
    int rc = make_map(&m2, stash, 4, d, s, box);
    if (rc) return rc;
  }
  WgtArgs a{};
  WgReduceArgs r{};
  int np = 0;
  const int ld0 = S + C + P;
  auto set_b = [&](WgtProblem& p, int layer, bool gh) {
    const int base = layer * kDgSlots * H;
    p.b_feat[0] = base;
    p.b_feat[1] = base + H;
    p.b_feat[2] = base + (gh ? 3 * H : 2 * H);  // (r, u, n_hh) or (r, u, n)
  };
  auto hcat = [&](WgtProblem& p, int tshift) {
    p.a_ctx = 0;
    p.a_cols[0] = kStashH * H;
    p.a_cols[1] = (NL > 1 ? kStashSlots + kStashH : kStashH) * H;
    p.a_tshift = tshift;
  };
  for (int c0 = 0; c0 < C; c0 += 128) {  // dW_ih0[:, S + c0 : S + c0 + 128]
    WgtProblem& p = a.prob[np];
    p.a_ctx = 1;
    for (int sl = 0; sl < 4; ++sl) p.a_cols[sl] = c0 + sl * 32;
    p.a_tshift = 0;
    set_b(p, 0, false);
    r.sc[np] = WgScatter{gw->w_ih[0], ld0, S + c0, 0, 128};
    ++np;
  }
  {  // dW_hh0 = d_gh0^T h0(t-1)
    WgtProblem& p = a.prob[np];
    hcat(p, -1);
    set_b(p, 0, true);
    r.sc[np] = WgScatter{gw->w_hh[0], H, 0, 0, H};
    ++np;
  }
  if (NL > 1) {
    {  // dW_ih1 = d_gi1^T h0(t)
      WgtProblem& p = a.prob[np];
      hcat(p, 0);
      set_b(p, 1, false);
      r.sc[np] = WgScatter{gw->w_ih[1], H, 0, 0, H};
      ++np;
    }
    {  // dW_hh1 = d_gh1^T h1(t-1)
      WgtProblem& p = a.prob[np];
      hcat(p, -1);
      set_b(p, 1, true);
      r.sc[np] = WgScatter{gw->w_hh[1], H, 0, H, H};
      ++np;
    }
  }
  a.nprob = np;
  a.nsplit = 148 / np;
  a.T = (int)T;
  a.total_kblocks = ntile * T * 4;
  if (a.total_kblocks < a.nsplit) a.nsplit = (int)a.total_kblocks;
  a.partials = partials;
  if ((size_t)np * a.nsplit * 128 * 192 > partial_floats) {
    set_error("tc_wgrads_tiled: workspace too small");
    return VISDE_EWORKSPACE;
  }
  const size_t smem = kWgtSmemBytes;
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(tc_wgrad_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  tc_wgrad_tiled_kernel<<<dim3(a.nsplit, np), kTcThreads, smem, st>>>(m0, m1, m2, a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  r.partials = partials;
  r.nsplit = a.nsplit;
  r.nprob = np;
  tc_wgrad_reduce_kernel<<<dim3((128 * 192 + 255) / 256, np), 256, 0, st>>>(r);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace visde
