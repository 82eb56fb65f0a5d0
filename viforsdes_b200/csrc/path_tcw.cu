// Tensor-core recurrence for WIDE state spaces (4 < S <= 10, H = 64, NL = 2): K1w path_fwd_tcw -- the forward of
// BASELINE config 5 (10-D Lorenz-96, 8192 trajectories) on tcgen05.
//
// Same tiling and numerics as path_tc.cu (128 trajectories per CTA = MMA M = TMEM lanes, fp16 hi/lo 3-pass products with
// FP32 accumulation, gate epilogue on all 8 warps, row-fastest tiled gi_ctx / stash).  What a wide state space changes:
//  * shared memory: W_out grows to 65 rows (mu + 55 Cholesky entries) and the state columns of W_ih_l0 to 30 per unit, so
//    the three recurrent matrices no longer fit next to the operand tiles.  W_hh_l0 stays resident; W_ih_l1 and W_hh_l1
//    TIME-SHARE one 48 KB buffer X: their swizzled hi/lo images (prepared once per launch by tcw_images_kernel) are
//    streamed in with cp.async.bulk behind the MMAs that read the previous occupant (commit -> mbarrier -> copy).
//  * TMEM: D0 192 + D1 256 columns leave 64; the output projection needs 65.  It is issued as two MMAs: the Cholesky rows
//    (N = 64) into the columns of D1's n_i block, which is dead between the layer-1 epilogue of step t and the layer-1 input
//    product of step t+1, and mu (N = 16) into the free columns.
//  * per-row S-sized I/O (eps in; mu, L, z out: 185 floats per trajectory-step) would be 32 scattered lines per warp access in
//    the [B,T,*] layouts, so eps is pre-tiled ([tile][t][S][128]) and the outputs are written as ONE row-fastest record per
//    step (mu | raw Cholesky | z_{t+1}); tcw_expand_kernel turns the records into paths / means / chol afterwards, the
//    backward reads them directly.
#include "path_tc.cuh"

namespace visde {
namespace {

constexpr int kFwdThreads = 256;

// Phase timestamps of one CTA (tools/tcw_trace.py; build with NVCCFLAGS+=-DVISDE_TCW_TRACE): thread 0 and thread 224 (warp 7) of
// CTA 0 write clock64() at the phase boundaries of steps 40..55 into a device array read back by visde_debug_tcw_trace
#ifdef VISDE_TCW_TRACE
__device__ long long g_tcw_trace_fwd[2 * 16 * 16];
#define TCW_TRACE(slot)                                                                             \
  do {                                                                                              \
    if (blockIdx.x == 0 && (tid == 0 || tid == 224) && t >= 40 && t < 56)                           \
      g_tcw_trace_fwd[((tid ? 1 : 0) * 16 + (int)(t - 40)) * 16 + (slot)] = clock64();              \
  } while (0)
#else
#define TCW_TRACE(slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------------
// weight images: one CTA per image
// ---------------------------------------------------------------------------------------------------------------------
// half_tiles: the images serve the 64-row kernels (path_*_tcw_kernel<S, 64>), whose two lane-half accumulators each hold 32 hidden
// units: forward tile rows are ordered [unit half][gate][32 units] so that the B rows of one half are contiguous
__global__ void __launch_bounds__(1024) tcw_images_kernel(PathParams p, uint8_t* __restrict__ img, int want_fwd, int want_bwd,
                                                          int half_tiles) {
  __shared__ uint32_t amax[3];
  const int tid = threadIdx.x, lane = tid & 31;
  const int S = p.S, NTRIL = p.n_tril, ld0 = p.S + p.C + p.P;
  if (tid == 0) amax[0] = amax[1] = amax[2] = 0u;
  __syncthreads();
  {
    float mx = 0.f, mo = 0.f, mz = 0.f;
    for (int idx = tid; idx < 192 * S; idx += blockDim.x) mz = fmaxf(mz, fabsf(p.w_ih[0][(int64_t)(idx / S) * ld0 + idx % S]));
    for (int m = 0; m < 3; ++m) {
      const float* src = m == 0 ? p.w_hh[0] : (m == 1 ? p.w_ih[1] : p.w_hh[1]);
      for (int idx = tid; idx < 192 * 64; idx += blockDim.x) mx = fmaxf(mx, fabsf(src[idx]));
    }
    for (int idx = tid; idx < p.n_out * 64; idx += blockDim.x) mo = fmaxf(mo, fabsf(p.out_w[idx]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mo = fmaxf(mo, __shfl_xor_sync(0xffffffffu, mo, o));
      mz = fmaxf(mz, __shfl_xor_sync(0xffffffffu, mz, o));
    }
    if (lane == 0) {
      atomicMax(&amax[0], __float_as_uint(mx));
      atomicMax(&amax[1], __float_as_uint(mo));
      atomicMax(&amax[2], __float_as_uint(mz));
    }
  }
  __syncthreads();
  const int ew = scale_exp(amax[0]), eo = scale_exp(amax[1]), ez = scale_exp(amax[2]);
  const float w_scale = exp2i(ew), o_scale = exp2i(eo), z_scale = exp2i(ez);
  const int job = blockIdx.x;  // 0..2 forward images, 3 W_out image, 4..6 backward images, 7 backward W_out, 8 backward W_z
  if (job == 0 && tid == 0) {
    reinterpret_cast<int*>(img)[0] = ew;
    reinterpret_cast<int*>(img)[1] = eo;
    reinterpret_cast<int*>(img)[2] = ez;
  }
  if (job < 3) {
    if (!want_fwd) return;
    const float* src = job == 0 ? p.w_hh[0] : (job == 1 ? p.w_ih[1] : p.w_hh[1]);
    uint8_t* thi = img + kImgFwd0 + (size_t)job * kWImg;
    uint8_t* tlo = thi + kWTileBytes;
    for (int idx = tid; idx < 192 * 8; idx += blockDim.x) {
      const int n = idx >> 3, c = idx & 7;
      const int srow = half_tiles ? ((n % 96) / 32) * 64 + (n / 96) * 32 + n % 32 : n;  // gate row of the nn.GRU matrix
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = src[srow * 64 + c * 8 + q] * w_scale;
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(thi + sw128(n, c)) = hi;
      *reinterpret_cast<uint4*>(tlo + sw128(n, c)) = lo;
    }
  } else if (job == 3) {
    if (!want_fwd) return;
    uint8_t* thi = img + kImgOut;
    uint8_t* tlo = thi + kOutRows * 128;
    for (int idx = tid; idx < kOutRows * 8; idx += blockDim.x) {
      const int n = idx >> 3, c = idx & 7;
      // tile row n: Cholesky entry n (out_w row S + n) for n < NTRIL, mu component n - 64 (out_w row n - 64) for n >= 64
      const int srow = n < NTRIL ? S + n : (n >= 64 && n - 64 < S ? n - 64 : -1);
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = srow >= 0 ? p.out_w[srow * 64 + c * 8 + q] * o_scale : 0.f;
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(thi + sw128(n, c)) = hi;
      *reinterpret_cast<uint4*>(tlo + sw128(n, c)) = lo;
    }
  } else if (job == 8) {
    if (!want_bwd) return;
    // state columns of W_ih_l0 as the B operand of d z_t (+)= d_gi_l0 . W_z: rows = state dim s (16, zero beyond S), K in the
    // (chunk, gate, unit) permutation of the transposed recurrent matrices
    uint8_t* thi = img + kImgWzBwd;
    uint8_t* tlo = thi + 3 * 16 * 128;
    for (int idx = tid; idx < 16 * 24; idx += blockDim.x) {
      const int sdim = idx & 15, hg = idx >> 4;
      const int gB = hg >> 1, half = hg & 1;
      // K-group gB: (chunk of 16 units, gate), or in the 64-row form (chunk of 32 units, gate, 16-unit half of the chunk)
      const int g = half_tiles ? (gB % 6) / 2 : gB % 3;
      const int u8 = half_tiles ? (gB / 6) * 32 + (gB % 2) * 16 + half * 8 : (gB / 3) * 16 + half * 8;
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = sdim < S ? p.w_ih[0][(int64_t)(g * 64 + u8 + q) * ld0 + sdim] * z_scale : 0.f;
      uint4 hi, lo;
      split8(x, hi, lo);
      const uint32_t off = (uint32_t)(gB >> 2) * 2048u + sw128(sdim, (gB & 3) * 2 + half);
      *reinterpret_cast<uint4*>(thi + off) = hi;
      *reinterpret_cast<uint4*>(tlo + off) = lo;
    }
  } else if (job == 7) {
    if (!want_bwd) return;
    // W_out as the B operand of dh_top (+)= d_out . W_out: rows = hidden unit i, K = output entries in the order of
    // kImgOutBwd (path_tc.cuh)
    uint8_t* thi = img + kImgOutBwd;
    uint8_t* tlo = thi + 64 * 128;
    for (int idx = tid; idx < 64 * 8; idx += blockDim.x) {
      const int i = idx >> 3, c = idx & 7;
      const int unit = half_tiles ? ((i & 31) >> 4) * 32 + (i >> 5) * 16 + (i & 15) : i;  // 64-row form: rows [unit half][chunk][16]
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = c * 8 + q;
        const int srow = k < NTRIL ? S + k : (k - NTRIL < S ? k - NTRIL : -1);
        x[q] = srow >= 0 ? p.out_w[srow * 64 + unit] * o_scale : 0.f;
      }
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(thi + sw128(i, c)) = hi;
      *reinterpret_cast<uint4*>(tlo + sw128(i, c)) = lo;
    }
  } else {
    if (!want_bwd) return;
    // transposed tiles of the backward: B operand rows = input / hidden index i (N = 64); K runs over (chunk c of 16 units,
    // gate g, unit in chunk): K-group gB = c * 3 + g sits in K-block gB / 4 at byte column (gB % 4) * 32 (path_tc_bwd.cu)
    const int m = job - 4;
    const float* src = m == 0 ? p.w_hh[0] : (m == 1 ? p.w_ih[1] : p.w_hh[1]);
    uint8_t* thi = img + kImgBwd0 + (size_t)m * kWImg;
    uint8_t* tlo = thi + kWTileBytes;
    for (int idx = tid; idx < 64 * 24; idx += blockDim.x) {
      const int i = idx & 63, hg = idx >> 6;
      const int gB = hg >> 1, half = hg & 1;
      // 64-row form: K-group gB = (chunk of 32 units, gate, 16-unit half of the chunk); B row i = output unit in [unit half][chunk][16]
      const int g = half_tiles ? (gB % 6) / 2 : gB % 3;
      const int u8 = half_tiles ? (gB / 6) * 32 + (gB % 2) * 16 + half * 8 : (gB / 3) * 16 + half * 8;
      const int unit = half_tiles ? ((i & 31) >> 4) * 32 + (i >> 5) * 16 + (i & 15) : i;
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = src[(g * 64 + u8 + q) * 64 + unit] * w_scale;
      uint4 hi, lo;
      split8(x, hi, lo);
      const uint32_t off = (uint32_t)(gB >> 2) * 8192u + sw128(i, (gB & 3) * 2 + half);
      *reinterpret_cast<uint4*>(thi + off) = hi;
      *reinterpret_cast<uint4*>(tlo + off) = lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// [B, T, F] rows (strided) -> features [f_off, f_off + F) of the row-fastest tiled record [tile][t][FD][128]
// ---------------------------------------------------------------------------------------------------------------------
// up to four sources per launch (the backward tiles gP | gM | gL | eps into one record); `steps` grid steps per block
struct TileSrc {
  const float* src;
  int64_t bstride, tstride;
  int F, f_off;
};
struct TileArgs {
  TileSrc s[4];
  int nsrc, FD, steps;
  int64_t B, T;
  float* dst;
};
__global__ void __launch_bounds__(256) tcw_tile_kernel(TileArgs a) {
  extern __shared__ float tile_s[];  // [128 rows][steps * Fmax + 1]
  const int64_t tb = blockIdx.y, t0 = (int64_t)blockIdx.x * a.steps;
  const int nt = (int)(a.T - t0 < a.steps ? a.T - t0 : a.steps);
  for (int q = 0; q < a.nsrc; ++q) {
    const TileSrc& sc = a.s[q];
    const int F = sc.F, pitch = a.steps * F + 1;
    const bool dense = sc.tstride == F;  // the nt * F floats of a row are contiguous
    if (q > 0) __syncthreads();
    // a dense row segment that starts on a 16-byte boundary and is a whole number of float4 is read with LDG.128
    const bool vec = dense && ((nt * F) & 3) == 0 && nt * F <= 256 && (sc.bstride & 3) == 0 && ((t0 * sc.tstride) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(sc.src) & 15u) == 0;
    if (vec) {
      // four rows per trip, their LDG.128 requested as one batch (volatile: not re-serialised behind the shared-memory stores)
      const int n4 = nt * F / 4, lane = threadIdx.x & 31;
      for (int r0 = threadIdx.x >> 5; r0 < kTileRows; r0 += 32) {
        float4 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t b = tb * kTileRows + r0 + 8 * u;
          const float4* row4 = reinterpret_cast<const float4*>(sc.src + (b < a.B ? b : a.B - 1) * sc.bstride + t0 * sc.tstride);
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const int e4 = lane + 32 * w;
            v[u][w] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e4 < n4)
              asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(v[u][w].x), "=f"(v[u][w].y), "=f"(v[u][w].z), "=f"(v[u][w].w)
                           : "l"(row4 + e4));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + 8 * u;
          const bool live = tb * kTileRows + r < a.B;  // pad rows carry exact zeros
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const int e4 = lane + 32 * w;
            if (e4 < n4) {
              float* d = tile_s + r * pitch + 4 * e4;
              d[0] = live ? v[u][w].x : 0.f; d[1] = live ? v[u][w].y : 0.f; d[2] = live ? v[u][w].z : 0.f; d[3] = live ? v[u][w].w : 0.f;
            }
          }
        }
      }
    } else {
      for (int r = threadIdx.x >> 5; r < kTileRows; r += 8) {
        const int64_t b = tb * kTileRows + r;
        for (int e = threadIdx.x & 31; e < nt * F; e += 32) {
          float v = 0.f;  // pad rows carry exact zeros
          if (b < a.B)
            v = dense ? sc.src[b * sc.bstride + t0 * sc.tstride + e] : sc.src[b * sc.bstride + (t0 + e / F) * sc.tstride + e % F];
          tile_s[r * pitch + e] = v;
        }
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nt * F * kTileRows; idx += 256) {
      const int r = idx & 127, e = idx >> 7;  // e = tt * F + f
      const int tt = e / F, f = e - tt * F;
      a.dst[((tb * a.T + t0 + tt) * a.FD + sc.f_off + f) * (int64_t)kTileRows + r] = tile_s[r * pitch + e];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward recurrence
// ---------------------------------------------------------------------------------------------------------------------
struct TcwFwdSmem {
  static constexpr int CS = 32;                            // floats per unit in c0: 3 S state weights + b_hn
  static constexpr int OFF_W0 = 0;                         // W_hh_l0 hi, lo
  static constexpr int OFF_X = kWImg;                      // time-shared: W_ih_l1 / W_hh_l1
  static constexpr int OFF_WOUT = 2 * kWImg;               // [hi, lo][80][128 B]
  static constexpr int OFF_A = OFF_WOUT + kOutImg;         // [2 layers][hi, lo][128][128 B]
  static constexpr int OFF_C0 = OFF_A + 4 * kATileBytes;   // float [64][CS]
  static constexpr int OFF_C1 = OFF_C0 + 64 * CS * 4;      // float [64][4]
  static constexpr int OFF_OUTB = OFF_C1 + 64 * 4 * 4;     // float [80]: bias per W_out tile row
  static constexpr int OFF_BAR = OFF_OUTB + kOutRows * 4;
  struct Bars {
    uint64_t init, d0, a0, d1, a1, out, wx, xr, pro;
    uint32_t tmem_base;
  };
  static constexpr size_t bytes = OFF_BAR + sizeof(Bars) + 1024;
};
static_assert(TcwFwdSmem::OFF_A % 1024 == 0 && TcwFwdSmem::OFF_WOUT % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(TcwFwdSmem::bytes <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void issue_gemm_w(uint32_t dcol, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                             uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint64_t dah = umma_desc(a_hi + j * 32, 16, 1024, 2), dal = umma_desc(a_lo + j * 32, 16, 1024, 2);
    const uint64_t dbh = umma_desc(b_hi + j * 32, 16, 1024, 2), dbl = umma_desc(b_lo + j * 32, 16, 1024, 2);
    umma_f16(dcol, dal, dbh, idesc, (accumulate || j > 0) ? 1u : 0u);
    umma_f16(dcol, dah, dbl, idesc, 1u);
    umma_f16(dcol, dah, dbh, idesc, 1u);
  }
}

// MT = trajectories per CTA = MMA M.  128: one CTA per tile, every TMEM lane a trajectory, two threads per row (32 hidden units each).
// 64: TWO CTAs per 128-row tile of the global layouts (sub-tile = work item & 1).  An M = 64 accumulator occupies lanes 0..15 of each
// 32-lane TMEM quadrant; every product is issued TWICE with N halved -- hidden units 0..31 into lanes 0..15, units 32..63 into the
// same columns of lanes 16..31 (D address + 16 lanes) -- so lane l of a warp owns row (l & 15) and unit half (l >> 4): four threads
// per row, 16 hidden units each, every lane busy.  The epilogues are issue-bound, so halving the rows per SM halves their time;
// used while the batch has at most SMs / 2 tiles.
template <int S, int MT>
__global__ void __launch_bounds__(kFwdThreads, 1) path_fwd_tcw_kernel(PathParams p) {
  using L = TcwFwdSmem;
  static_assert(MT == 128 || MT == 64, "MMA M");
  constexpr int LPQ = MT / 4, SUBS = kTileRows / MT;
  constexpr bool HALF = MT == 64;
  constexpr int UPT = HALF ? 16 : 32;  // hidden units per thread
  constexpr int HU = HALF ? 32 : 64;   // hidden units per accumulator block (one lane half)
  constexpr int NCH = UPT / 8, GA = UPT / 2;
  constexpr int NTRIL = S * (S + 1) / 2, CS = L::CS, OF = tcw_out_feats(S);
  static_assert(S > 4 && S <= kTcwMaxS && NTRIL <= 64 && 3 * S + 1 <= CS, "wide-state tensor-core recurrence: 4 < S <= 10");
  constexpr uint32_t TMEM_COLS = 512;
  // M = 128: D0 0..191 | D1 192..447 (r, u, n_i, n_h) | Cholesky rows alias D1's n_i block | mu 448..463
  // M = 64 : D0 0..95  | D1 96..223                    | Cholesky rows 224..287            | mu 288..303 (both lane halves)
  constexpr uint32_t D0_COL = 0, D1_COL = 3 * HU, TRIL_COL = HALF ? 7 * HU : D1_COL + 2 * HU, MU_COL = HALF ? 7 * HU + 64 : 448;
  extern __shared__ __align__(1024) uint8_t smem_raw_tcw[];
  uint8_t* smem = smem_raw_tcw + ((1024u - (smem_u32(smem_raw_tcw) & 1023u)) & 1023u);
  typename L::Bars* bars = reinterpret_cast<typename L::Bars*>(smem + L::OFF_BAR);
  float* c0 = reinterpret_cast<float*>(smem + L::OFF_C0);
  float* c1 = reinterpret_cast<float*>(smem + L::OFF_C1);
  float* outb = reinterpret_cast<float*>(smem + L::OFF_OUTB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld0 = S + p.C + p.P;
  const int64_t T = p.T;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.wimg);
  const int ew = reinterpret_cast<const int*>(img)[0], eo = reinterpret_cast<const int*>(img)[1];
  const float sc = exp2i(-(ew + kHExp)), sco = exp2i(-(eo + kHExp));

  for (int idx = tid; idx < 64 * CS; idx += kFwdThreads) {  // per unit: (r_s, u_s) pairs, then n_s, then b_hn
    const int j = idx / CS, q = idx % CS;
    float v = 0.f;
    if (q < 2 * S) v = p.w_ih[0][(int64_t)((q & 1) * 64 + j) * ld0 + (q >> 1)];
    else if (q < 3 * S) v = p.w_ih[0][(int64_t)(128 + j) * ld0 + (q - 2 * S)];
    else if (q == 3 * S) v = p.b_hh[0][128 + j];
    c0[idx] = v;
  }
  for (int j = tid; j < 64; j += kFwdThreads) {
    c1[j * 4 + 0] = p.b_ih[1][j] + p.b_hh[1][j];
    c1[j * 4 + 1] = p.b_ih[1][64 + j] + p.b_hh[1][64 + j];
    c1[j * 4 + 2] = p.b_ih[1][128 + j];
    c1[j * 4 + 3] = p.b_hh[1][128 + j];
  }
  if (tid < kOutRows) outb[tid] = tid < NTRIL ? p.out_b[S + tid] : (tid >= 64 && tid - 64 < S ? p.out_b[tid - 64] : 0.f);

  if (tid == 0) {
    mbar_init(&bars->init, kEpiThreads);
    mbar_init(&bars->d0, MT == 64 ? 2 : 1);  // one commit per thread that issues into the barrier
    mbar_init(&bars->a0, kEpiThreads);
    mbar_init(&bars->d1, MT == 64 ? 4 : 1);
    mbar_init(&bars->a1, kEpiThreads);
    mbar_init(&bars->out, MT == 64 ? 4 : 1);
    mbar_init(&bars->wx, 1);
    mbar_init(&bars->xr, MT == 64 ? 2 : 1);
    mbar_init(&bars->pro, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const int64_t nitems = (p.B + kTileRows - 1) / kTileRows * SUBS;
  if (tid == 0) {
    // resident tiles on `pro`, the first occupant of X (W_ih_l1) on `wx`
    mbar_expect_tx(&bars->pro, kWImg + kOutImg);
    bulk_load_1d(smem + L::OFF_W0, img + kImgFwd0, kWImg, &bars->pro);
    bulk_load_1d(smem + L::OFF_WOUT, img + kImgOut, kOutImg, &bars->pro);
    mbar_expect_tx(&bars->wx, kWImg);
    bulk_load_1d(smem + L::OFF_X, img + kImgFwd0 + kWImg, kWImg, &bars->wx);
  }

  // ---- MMA issue (lane 0 of warp 0, between its epilogue phases) ---------------------------------
  const uint32_t w0h = smem_u32(smem + L::OFF_W0), w0l = w0h + kWTileBytes;
  const uint32_t xh = smem_u32(smem + L::OFF_X), xl = xh + kWTileBytes;
  const uint32_t a0h = smem_u32(smem + L::OFF_A), a0l = a0h + kATileBytes;
  const uint32_t a1h = a0h + 2 * kATileBytes, a1l = a1h + kATileBytes;
  const uint32_t woh = smem_u32(smem + L::OFF_WOUT), wol = woh + kOutRows * 128;
  constexpr uint32_t ID192 = idesc_f16(192, MT), ID128 = idesc_f16(128, MT), ID64 = idesc_f16(64, MT), ID16 = idesc_f16(16, MT);
  constexpr uint32_t ID96 = idesc_f16(96, MT), ID32 = idesc_f16(32, MT);
  constexpr uint32_t NROWS = 128 * 128;  // byte offset of gate rows 128.. (the n block) in a weight tile (M = 128 row order)
  constexpr uint32_t HROWS = 96 * 128;   // M = 64 row order [half][gate][32]: byte offset of the second unit half
  constexpr uint32_t LHALF = 16u << 16;  // TMEM address of the second lane half
  uint32_t ph_init = 0, ph_a0 = 0, ph_a1 = 0, ph_wx = 0, ph_xr = 0;  // issuer-side phases (thread 0)
  // M = 64: FOUR issuing threads side by side (lane 0 of warps 0..3): warp w issues unit half h = w & 1, and of the two products
  // on the critical path of a hand-off (layer-1 input: r, u | n_i; output: Cholesky rows | mu) block w >> 1; the two products that
  // have a whole phase to complete (h0 . W_hh_l0, h1 . W_hh_l1) ride with block 0.  A hand-off then costs the issue time of 12 MMAs.
  // Every completion barrier of the tensor pipe counts one commit per thread that issued into it.
  constexpr int NISS = HALF ? 4 : 1;
  constexpr uint32_t NI_OFF = HALF ? 3 * HU : 2 * HU, NH_OFF = HALF ? 2 * HU : 3 * HU;  // n_i / n_h blocks of D1
  const uint32_t hh = HALF ? (uint32_t)(warp & 1) : 0u;
  const bool blk0 = !HALF || (warp >> 1) == 0, blk1 = !HALF || (warp >> 1) == 1;
  const uint32_t tm_h = tmem + hh * LHALF, wro = hh * HROWS;
  auto issue_recurrent_l0 = [&]() {  // D0 = h0 . W_hh_l0^T (for the next step)
    if (!blk0) return;
    if (HALF) issue_gemm_w(tm_h + D0_COL, a0h, a0l, w0h + wro, w0l + wro, ID96, false);
    else issue_gemm_w(tmem + D0_COL, a0h, a0l, w0h, w0l, ID192, false);
    umma_commit(&bars->d0);
  };
  auto mma_x_recurrent_l1 = [&]() {  // X holds W_hh_l1: D1[r, u] = h1 . W_hh_l1[r, u]^T, D1[n_h] = h1 . W_hh_l1[n]^T
    if (!blk0) return;
    if (HALF) {  // r, u, n_h are adjacent: one N = 96 product
      issue_gemm_w(tm_h + D1_COL, a1h, a1l, xh + wro, xl + wro, ID96, false);
    } else {
      issue_gemm_w(tmem + D1_COL, a1h, a1l, xh, xl, ID128, false);
      issue_gemm_w(tmem + D1_COL + NH_OFF, a1h, a1l, xh + NROWS, xl + NROWS, ID64, false);
    }
  };
  auto issue_x_recurrent_l1 = [&]() {
    mma_x_recurrent_l1();
    if (blk0) umma_commit(&bars->xr);
  };
  auto issue_x_input_l1 = [&]() {  // X holds W_ih_l1: r, u accumulate onto the recurrent part, n_i has its own columns
    if (HALF) {
      if (blk0) issue_gemm_w(tm_h + D1_COL, a0h, a0l, xh + wro, xl + wro, ID64, true);
      else issue_gemm_w(tm_h + D1_COL + NI_OFF, a0h, a0l, xh + wro + 64 * 128, xl + wro + 64 * 128, ID32, false);
    } else {
      issue_gemm_w(tmem + D1_COL, a0h, a0l, xh, xl, ID128, true);
      issue_gemm_w(tmem + D1_COL + NI_OFF, a0h, a0l, xh + NROWS, xl + NROWS, ID64, false);
    }
    umma_commit(&bars->d1);
  };
  auto issue_out = [&]() {  // M = 128: Cholesky rows into the (dead) n_i columns of D1, mu into the free columns;
                            // M = 64: every lane half gets the whole output row in its own columns
    if (blk0) issue_gemm_w(tm_h + TRIL_COL, a1h, a1l, woh, wol, ID64, false);
    if (blk1) issue_gemm_w(tm_h + MU_COL, a1h, a1l, woh + 64 * 128, wol + 64 * 128, ID16, false);
    umma_commit(&bars->out);
  };
  auto load_x = [&](int m) {  // thread 0: X <- image of W_ih_l1 (m = 1) / W_hh_l1 (m = 2); the MMAs reading X have completed
    mbar_expect_tx(&bars->wx, kWImg);
    bulk_load_1d(smem + L::OFF_X, img + kImgFwd0 + (size_t)m * kWImg, kWImg, &bars->wx);
  };
  auto wait_x = [&]() {
    mbar_wait(&bars->wx, ph_wx);
    ph_wx ^= 1;
    tc_fence_after();
  };
  if (warp < NISS) mbar_wait(&bars->pro, 0);

  {
    const int quad = warp & 3, cg = warp >> 2;
    const int lh = HALF ? lane >> 4 : 0;              // M = 64: unit half = TMEM lane half of this thread
    const int row = quad * LPQ + (lane & (LPQ - 1));  // row of the CTA's A tile / accumulators
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    const int jc0 = cg * UPT;        // first of this thread's units inside its accumulator block (column offset)
    const int u0 = lh * HU + jc0;    // ... as a hidden-unit index
    uint8_t* a_tiles = smem + L::OFF_A;
    const float hs = exp2i(kHExp);
    uint32_t ph_d0 = 0, ph_d1 = 0, ph_out = 0;

    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int64_t tile = item / SUBS;
      const int grow = (int)(item % SUBS) * MT + row;  // row of the 128-row tile of the global layouts
      const int64_t b_raw = tile * kTileRows + grow;
      const bool ok = b_raw < p.B;
      const int64_t b = b_raw < p.B ? b_raw : p.B - 1;
      const bool writer = cg == 0 && lh == 0;  // pad rows write too: the tiled records of pad rows are never read as data
      // X holds W_ih_l1 at every tile start.  The first tile's copy was issued in the prologue; a later tile re-issues it so
      // that every tile consumes exactly one completion of `wx` at its step 0 (the last reader of X, the layer-1 input
      // product of the previous tile's final step, completed before this thread passed that step's d1 wait)
      if (tid == 0 && item != (int64_t)blockIdx.x) load_x(1);
      // h(-1) = 0
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const uint32_t off = sw128(row, (u0 >> 3) + c);
          *reinterpret_cast<uint4*>(a_tiles + (2 * k) * kATileBytes + off) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(a_tiles + (2 * k + 1) * kATileBytes + off) = make_uint4(0, 0, 0, 0);
        }
      fence_proxy_async();
      mbar_arrive(&bars->init);
      if (warp < NISS) {
        mbar_wait(&bars->init, ph_init);
        ph_init ^= 1;
        tc_fence_after();
        if (elect_one_sync()) {
          // h(-1) = 0: the A tiles are zero, so these just clear the accumulators (X holds the finite W_ih_l1 image: the
          // recurrent product of layer 1 is issued with whatever finite tile is there, 0 . w = 0)
          issue_recurrent_l0();
          mbar_wait(&bars->wx, ph_wx);  // no phase flip: issue_x_input_l1 of step 0 waits for the same completion
          tc_fence_after();
          mma_x_recurrent_l1();
        }
        __syncwarp();
      }

      float z[S], hprev[2][UPT];
#pragma unroll
      for (int s = 0; s < S; ++s) z[s] = ok ? p.x0[b * S + s] : 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int j = 0; j < UPT; ++j) hprev[k][j] = 0.f;
      const float* gi_p = p.gi_ctx + tile * T * (int64_t)(192 * kTileRows) + (int64_t)u0 * kTileRows + grow;
      float* st_p = p.stash ? p.stash + tile * T * (int64_t)(2 * kStashSlots * 64 * kTileRows) + (int64_t)u0 * kTileRows + grow
                                   : nullptr;
      const float* eps_p = p.epst + tile * T * (int64_t)(S * kTileRows) + grow;
      float* ot_p = p.otile + tile * T * (int64_t)(OF * kTileRows) + grow;

      float g01[3][GA], g23[3][GA];  // context part of the layer-0 gates: first / second half of this thread's units
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int q = 0; q < GA; ++q) g01[g][q] = gi_p[(g * 64 + q) * kTileRows];

      for (int64_t t = 0; t < T; ++t) {
        const bool has_next = t + 1 < T;
        TCW_TRACE(0);
        // ---------------- layer 0 ----------------
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int q = 0; q < GA; ++q) g23[g][q] = gi_p[(g * 64 + GA + q) * kTileRows];
        mbar_wait(&bars->d0, ph_d0);
        ph_d0 ^= 1;
        tc_fence_after();
        TCW_TRACE(1);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int j0 = u0 + c * 8, jc = jc0 + c * 8;
          uint32_t dr[8], du[8], dn[8];
          tmem_ld8_nowait(tl + D0_COL + jc, dr);
          tmem_ld8_nowait(tl + D0_COL + HU + jc, du);
          tmem_ld8_nowait(tl + D0_COL + 2 * HU + jc, dn);
          float gcur[3][8];
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int q = 0; q < 8; ++q)
              gcur[g][q] = c < NCH / 2 ? g01[g][(c % (NCH / 2)) * 8 + q] : g23[g][(c % (NCH / 2)) * 8 + q];
          tmem_ld_wait();
          float hx[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int j = j0 + jj;
            const float* cw = c0 + j * CS;
            float2 pru = make_float2(fmaf(sc, __uint_as_float(dr[jj]), gcur[0][jj]), fmaf(sc, __uint_as_float(du[jj]), gcur[1][jj]));
            float pni = gcur[2][jj];
            // state columns of W_ih_l0: 3 S weights per unit, broadcast reads (every lane of the warp has the same unit);
            // the r and u rows are interleaved so that one packed FFMA2 serves both gates
            float cc[CS];
#pragma unroll
            for (int v = 0; v < CS / 4; ++v) {
              const float4 w4 = *reinterpret_cast<const float4*>(cw + 4 * v);
              cc[4 * v] = w4.x; cc[4 * v + 1] = w4.y; cc[4 * v + 2] = w4.z; cc[4 * v + 3] = w4.w;
            }
            const float pnh = fmaf(sc, __uint_as_float(dn[jj]), cc[3 * S]);
#pragma unroll
            for (int s = 0; s < S; ++s) {
              fma2(pru, make_float2(cc[2 * s], cc[2 * s + 1]), make_float2(z[s], z[s]));
              pni = fmaf(cc[2 * S + s], z[s], pni);
            }
            const float pr = pru.x, pu = pru.y;
            const float r = sigmoid_f(pr);
            const float n = tanh_f(fmaf(r, pnh, pni));
            const float u = sigmoid_f(pu);
            const float hn = fmaf(u, hprev[0][c * 8 + jj] - n, n);
            hprev[0][c * 8 + jj] = hn;
            hx[jj] = hn * hs;
            if (st_p) {
              float* st = st_p + (c * 8 + jj) * kTileRows;
              st[kStashR * 64 * kTileRows] = r;
              st[kStashU * 64 * kTileRows] = u;
              st[kStashN * 64 * kTileRows] = n;
              st[kStashNhh * 64 * kTileRows] = pnh;
              st[kStashH * 64 * kTileRows] = hn;
            }
          }
          uint4 hi, lo;
          split8(hx, hi, lo);
          const uint32_t off = sw128(row, j0 >> 3);
          *reinterpret_cast<uint4*>(a_tiles + off) = hi;
          *reinterpret_cast<uint4*>(a_tiles + kATileBytes + off) = lo;
        }
        fence_proxy_async();
        tc_fence_before();
        TCW_TRACE(2);
        mbar_arrive(&bars->a0);
        gi_p += 192 * kTileRows;
        if (warp < NISS) {
          mbar_wait(&bars->a0, ph_a0);
          ph_a0 ^= 1;
          tc_fence_after();
          if (elect_one_sync()) {
            wait_x();  // X = W_ih_l1
            issue_x_input_l1();
            if (has_next) issue_recurrent_l0();
          }
          __syncwarp();
        }
        if (has_next) {
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int q = 0; q < GA; ++q) g01[g][q] = gi_p[(g * 64 + q) * kTileRows];
        }
        // this step's noise (tiled: one coalesced line per component), in flight during layer 1
        float eps[S];
#pragma unroll
        for (int s = 0; s < S; ++s) eps[s] = eps_p[s * kTileRows];
        eps_p += S * kTileRows;

        // ---------------- layer 1 ----------------
        TCW_TRACE(3);
        mbar_wait(&bars->d1, ph_d1);
        ph_d1 ^= 1;
        tc_fence_after();
        TCW_TRACE(4);
        if (tid == 0 && has_next) load_x(2);  // the layer-1 input product has finished reading X: stream W_hh_l1 in
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int j0 = u0 + c * 8, jc = jc0 + c * 8;
          uint32_t dr[8], du[8], di[8], dn[8];
          tmem_ld8_nowait(tl + D1_COL + jc, dr);
          tmem_ld8_nowait(tl + D1_COL + HU + jc, du);
          tmem_ld8_nowait(tl + D1_COL + NI_OFF + jc, di);
          tmem_ld8_nowait(tl + D1_COL + NH_OFF + jc, dn);
          tmem_ld_wait();
          float hx[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int j = j0 + jj;
            const float4 cb = *reinterpret_cast<const float4*>(c1 + j * 4);
            const float pr = fmaf(sc, __uint_as_float(dr[jj]), cb.x);
            const float pu = fmaf(sc, __uint_as_float(du[jj]), cb.y);
            const float pni = fmaf(sc, __uint_as_float(di[jj]), cb.z);
            const float pnh = fmaf(sc, __uint_as_float(dn[jj]), cb.w);
            const float r = sigmoid_f(pr);
            const float n = tanh_f(fmaf(r, pnh, pni));
            const float u = sigmoid_f(pu);
            const float hn = fmaf(u, hprev[1][c * 8 + jj] - n, n);
            hprev[1][c * 8 + jj] = hn;
            hx[jj] = hn * hs;
            if (st_p) {
              float* st = st_p + (kStashSlots * 64 + c * 8 + jj) * kTileRows;
              st[kStashR * 64 * kTileRows] = r;
              st[kStashU * 64 * kTileRows] = u;
              st[kStashN * 64 * kTileRows] = n;
              st[kStashNhh * 64 * kTileRows] = pnh;
              st[kStashH * 64 * kTileRows] = hn;
            }
          }
          uint4 hi, lo;
          split8(hx, hi, lo);
          const uint32_t off = sw128(row, j0 >> 3);
          *reinterpret_cast<uint4*>(a_tiles + 2 * kATileBytes + off) = hi;
          *reinterpret_cast<uint4*>(a_tiles + 3 * kATileBytes + off) = lo;
        }
        fence_proxy_async();
        tc_fence_before();
        TCW_TRACE(5);
        mbar_arrive(&bars->a1);
        if (warp < NISS) {
          mbar_wait(&bars->a1, ph_a1);
          ph_a1 ^= 1;
          tc_fence_after();
          if (elect_one_sync()) {
            issue_out();
            if (has_next) {
              wait_x();  // X = W_hh_l1
              issue_x_recurrent_l1();
            }
          }
          __syncwarp();
        }
        if (st_p) st_p += 2 * kStashSlots * 64 * kTileRows;

        // ---------------- output projection + reparameterised Euler-Maruyama update ----------------
        TCW_TRACE(6);
        mbar_wait(&bars->out, ph_out);
        ph_out ^= 1;
        tc_fence_after();
        TCW_TRACE(7);
        float mu[S], acc[S];
        {
          uint32_t mv[16];
          tmem_ld16_nowait(tl + MU_COL, mv);
          tmem_ld_wait();
#pragma unroll
          for (int s = 0; s < S; ++s) {
            mu[s] = fmaf(sco, __uint_as_float(mv[s]), outb[64 + s]);
            acc[s] = 0.f;
            if (writer) ot_p[s * kTileRows] = mu[s];
          }
        }
        {
          uint32_t lv[8];
#pragma unroll
          for (int rr = 0; rr < S; ++rr) {
#pragma unroll
            for (int jj = 0; jj <= rr; ++jj) {
              const int ti = rr * (rr + 1) / 2 + jj;  // row-major: ascending, so the 8-column groups are loaded in order
              if (ti % 8 == 0) {
                tmem_ld8_nowait(tl + TRIL_COL + ti, lv);
                tmem_ld_wait();
              }
              const float raw = fmaf(sco, __uint_as_float(lv[ti % 8]), outb[ti]);
              const float Lv = (jj == rr) ? fmaxf(raw, VISDE_DIAG_MIN) : raw;
              acc[rr] = fmaf(Lv, eps[jj], acc[rr]);
              if (writer) ot_p[(S + ti) * kTileRows] = raw;
            }
          }
        }
        tc_fence_before();
#pragma unroll
        for (int s = 0; s < S; ++s) {
          z[s] = z[s] + mu[s] * p.dt + acc[s] * p.sqrt_dt;
          if (writer) ot_p[(S + NTRIL + s) * kTileRows] = z[s];
        }
        ot_p += OF * kTileRows;
        TCW_TRACE(8);
        if (tid == 0 && has_next) {
          // the recurrent product of layer 1 has finished reading X: stream W_ih_l1 back in for the next step
          mbar_wait(&bars->xr, ph_xr);
          ph_xr ^= 1;
          load_x(1);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// tiled step records -> paths [B,T+1,S], means [B,T,S], chol [B,T,S,S] (floored diagonal, zero upper triangle)
// ---------------------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256) tcw_expand_kernel(const float* __restrict__ otile, const float* __restrict__ x0, int64_t B,
                                                         int64_t T, float* __restrict__ paths, float* __restrict__ means,
                                                         float* __restrict__ chol) {
  constexpr int NTRIL = S * (S + 1) / 2, OF = tcw_out_feats(S), PITCH = kTileRows + 1;
  __shared__ float rec[OF * PITCH];
  const int64_t t = blockIdx.x, tb = blockIdx.y;
  const float* src = otile + (tb * T + t) * (int64_t)(OF * kTileRows);
  for (int idx = threadIdx.x; idx < OF * kTileRows; idx += 256) rec[(idx >> 7) * PITCH + (idx & 127)] = src[idx];
  __syncthreads();
  const int64_t b0 = tb * kTileRows;
  const int nrow = (int)(B - b0 < kTileRows ? B - b0 : kTileRows);
  for (int idx = threadIdx.x; idx < nrow * S; idx += 256) {
    const int r = idx / S, s = idx - r * S;
    means[((b0 + r) * T + t) * S + s] = rec[s * PITCH + r];
    paths[((b0 + r) * (T + 1) + t + 1) * S + s] = rec[(S + NTRIL + s) * PITCH + r];
    if (t == 0) paths[(b0 + r) * (T + 1) * S + s] = x0[(b0 + r) * S + s];
  }
  for (int idx = threadIdx.x; idx < nrow * S * S; idx += 256) {
    const int r = idx / (S * S), e = idx - r * (S * S), i = e / S, j = e - i * S;
    float v = 0.f;
    if (j <= i) {
      v = rec[(S + i * (i + 1) / 2 + j) * PITCH + r];
      if (j == i) v = fmaxf(v, VISDE_DIAG_MIN);
    }
    chol[((b0 + r) * T + t) * (int64_t)(S * S) + e] = v;
  }
}

template <int S>
int launch_fwd_tcw(const PathParams& p, cudaStream_t st) {
  const size_t smem = TcwFwdSmem::bytes;
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_fwd_tcw_kernel<S, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_fwd_tcw_kernel<S, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (p.B + kTileRows - 1) / kTileRows;
  if (tcw_half_tiles(ntiles, sms))
    path_fwd_tcw_kernel<S, 64><<<(unsigned)(2 * ntiles), kFwdThreads, smem, st>>>(p);
  else
    path_fwd_tcw_kernel<S, 128><<<(unsigned)(ntiles < sms ? ntiles : sms), kFwdThreads, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  tcw_expand_kernel<S><<<dim3((unsigned)p.T, (unsigned)ntiles), 256, 0, st>>>(p.otile, p.x0, p.B, p.T, p.paths, p.means, p.chol);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace

#ifdef VISDE_TCW_TRACE
extern "C" int visde_debug_tcw_trace_fwd(long long* out) {
  return cudaMemcpyFromSymbol(out, g_tcw_trace_fwd, sizeof(g_tcw_trace_fwd)) == cudaSuccess ? 0 : -2;
}
#endif

bool tcw_rec_supported(const PathParams& p) {
  return p.H == 64 && p.NL == 2 && p.S > 4 && p.S <= kTcwMaxS && p.T >= 1 && p.T <= 65535 && p.B >= 1 &&
         p.T * (int64_t)(p.NL * kStashSlots * p.H) < (int64_t(1) << 31);
}

int launch_tcw_images(const PathParams& p, void* img, bool fwd, bool bwd, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int half_tiles = tcw_half_tiles((p.B + kTileRows - 1) / kTileRows, sms) ? 1 : 0;
  tcw_images_kernel<<<9, 1024, 0, st>>>(p, reinterpret_cast<uint8_t*>(img), fwd ? 1 : 0, bwd ? 1 : 0, half_tiles);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

// srcs: nsrc x (pointer, batch stride, time stride, F, feature offset in the record)
int launch_tcw_tile_multi(const float* const* src, const int64_t* bstride, const int64_t* tstride, const int* F, const int* f_off,
                          int nsrc, int64_t B, int64_t T, float* dst, int FD, cudaStream_t st) {
  TileArgs a{};
  int fmax = 1;
  for (int q = 0; q < nsrc; ++q) {
    a.s[q] = TileSrc{src[q], bstride[q], tstride[q], F[q], f_off[q]};
    if (F[q] > fmax) fmax = F[q];
  }
  a.nsrc = nsrc;
  a.FD = FD;
  a.steps = fmax >= 96 ? 1 : (192 / fmax > 16 ? 16 : 192 / fmax);
  a.B = B;
  a.T = T;
  a.dst = dst;
  const int64_t ntiles = (B + kTileRows - 1) / kTileRows;
  const size_t smem = sizeof(float) * kTileRows * (a.steps * fmax + 1);
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(tcw_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.done(attr_dev);
  }
  if (smem > 200 * 1024) {
    set_error("tile: %d features per step do not fit the staging buffer", fmax);
    return VISDE_EINVAL;
  }
  tcw_tile_kernel<<<dim3((unsigned)((T + a.steps - 1) / a.steps), (unsigned)ntiles), 256, smem, st>>>(a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int launch_tcw_tile(const float* src, int64_t B, int64_t T, int F, int64_t bstride, int64_t tstride, float* dst, int FD, int f_off,
                    cudaStream_t st) {
  return launch_tcw_tile_multi(&src, &bstride, &tstride, &F, &f_off, 1, B, T, dst, FD, st);
}

int launch_path_fwd_tcw(const PathParams& p, cudaStream_t st) {
  // eps [B,T,S] -> [tile][t][S][128]
  int rc = launch_tcw_tile(p.eps, p.B, p.T, p.S, p.T * (int64_t)p.S, p.S, p.epst, p.S, 0, st);
  if (rc) return rc;
  switch (p.S) {
    case 5: return launch_fwd_tcw<5>(p, st);
    case 6: return launch_fwd_tcw<6>(p, st);
    case 7: return launch_fwd_tcw<7>(p, st);
    case 8: return launch_fwd_tcw<8>(p, st);
    case 9: return launch_fwd_tcw<9>(p, st);
    case 10: return launch_fwd_tcw<10>(p, st);
  }
  set_error("wide-state tensor-core recurrence: unsupported state dim %d", p.S);
  return VISDE_EINVAL;
}

}  // namespace visde
