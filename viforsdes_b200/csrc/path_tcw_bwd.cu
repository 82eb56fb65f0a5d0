// Tensor-core BPTT for WIDE state spaces (4 < S <= 10, H = 64, NL = 2): K2w path_bwd_tcw -- the backward of BASELINE
// config 5 on tcgen05.  Reverse-time mirror of path_tcw.cu; MMA structure, row-scaled fp16 hi/lo chunks, ring and TMEM
// plan are those of path_tc_bwd.cu.  What the wide state space changes:
//  * shared memory: W_out^T (65 x 64 fp32) and the state columns of W_ih_l0 (64 x 30) no longer fit next to the three
//    transposed recurrent matrices.  W_ih_l1^T stays resident; W_hh_l1^T (layer-1 phase of a step) and W_hh_l0^T (layer-0
//    phase) TIME-SHARE one 48 KB buffer Y, streamed in from the launch's tile images with cp.async.bulk behind the MMAs
//    of the previous phase (in0 / empty commits -> copy -> wy).
//  * the cotangent of the output projection has 65 entries per trajectory-step: it is never held in registers.  Entry m
//    is formed from the tiled cotangent record (gP | gM | gL | eps, one coalesced line per value), written to the tiled
//    d_out buffer and immediately contracted with row m of W_out for this thread's 32 hidden units (65 x 32 FFMA from
//    broadcast shared-memory reads).
//  * the S-sized reductions over (b, t) (dW_ih_l0[:, :S], dW_out, db_out, biases, sum_t d_gi) are time-parallel passes over
//    the tiled d_pre / d_out / stash / step records (tcw_thin_* below), fixed-order sums: bit-deterministic.
#include <type_traits>

#include "path_tc.cuh"

namespace visde {
namespace {

constexpr int kBwdThreads = 256;
#ifdef VISDE_TCW_TRACE
__device__ long long g_tcw_trace_bwd[2 * 16 * 16];
#define TCWB_TRACE(slot)                                                                            \
  do {                                                                                              \
    if (blockIdx.x == 0 && (tid == 0 || tid == 224) && t >= 40 && t < 56)                           \
      g_tcw_trace_bwd[((tid ? 1 : 0) * 16 + (t - 40)) * 16 + (slot)] = clock64();                   \
  } while (0)
#else
#define TCWB_TRACE(slot) do { } while (0)
#endif
constexpr int kRowExp = 9;  // rows are scaled so that max|dh| 2^e is in [2^9, 2^10)

template <int S>
struct TcwBwdSmem {
  static constexpr int NOUT = S + S * (S + 1) / 2;
  static constexpr int CZ = 32;
  static constexpr int OFF_W1 = 0;                                // W_ih_l1^T hi, lo (resident)
  static constexpr int OFF_Y = kWImg;                             // time-shared: W_hh_l1^T / W_hh_l0^T
  static constexpr int OFF_A = 2 * kWImg;                         // ring [2][hi, lo][128][128 B]
  static constexpr int NTRIL = S * (S + 1) / 2;
  static constexpr int KMU = 64 - NTRIL < S ? 64 - NTRIL : S;     // mu components that ride in the K = 64 MMA operand
  static constexpr int NREST = S - KMU;                            // mu components contracted on the FP32 path
  static constexpr int OFF_WOUT = OFF_A + 4 * kATileBytes;        // W_out B tile [hi, lo][64 units][128 B]
  static constexpr int OFF_WZ = OFF_WOUT + kOutBwdImg;            // W_z B tile [hi, lo][3 K-blocks][16 state dims][128 B]
  static constexpr int OFF_WREST = OFF_WZ + kWzBwdImg;            // float [NREST][64]: W_out rows of the remaining mu components
  static constexpr int OFF_MAX = OFF_WREST + (NREST > 0 ? NREST : 1) * 64 * 4;  // float [2 buffers][2 cg][128]
  static constexpr int OFF_BAR = (OFF_MAX + 2 * 2 * 128 * 4 + 15) / 16 * 16;
  struct Bars {
    uint64_t full[2], empty[2], in0, wy, pro, outd, dzr;
    uint32_t tmem_base;
  };
  static constexpr size_t bytes = OFF_BAR + sizeof(Bars) + 1024;
};

__device__ __forceinline__ int row_exp_w(float mx) {
  const uint32_t bits = __float_as_uint(mx);
  if (bits == 0u) return 0;
  int e = kRowExp - ((int)(bits >> 23) - 127);
  e = e > 100 ? 100 : e;
  return e < -100 ? -100 : e;
}

// MT: trajectories per CTA = MMA M.  128: two threads per row, 4 chunks of 16 hidden units per layer phase (8 units per thread and
// chunk), accumulators 64 columns wide.  64 (two CTAs per tile of the global layouts, see path_fwd_tcw_kernel): FOUR threads per row
// -- lane l of a warp owns row l & 15 and, as lane half l >> 4, the output units whose accumulators sit on TMEM lanes 16 (l >> 4) ..
// of its quadrant: every product is issued twice with N = 32 (unit half h into D + 16 h lanes).  A layer phase has 2 chunks of 32
// hidden units; a chunk's A operand is two K = 64 tiles (gates r | u and n | n_hh) of 64 rows; lane half h owns units
// {32 c + 16 h + w}: B rows (output units) are ordered [h][c][16] and K runs (chunk, gate, 32 units) in the images of this form.
template <int S, int MT>
__global__ void __launch_bounds__(kBwdThreads, 1) path_bwd_tcw_kernel(PathParams p) {
  using L = TcwBwdSmem<S>;
  static_assert(MT == 128 || MT == 64, "MMA M");
  constexpr int LPQ = MT / 4, SUBS = kTileRows / MT;
  constexpr bool HALF = MT == 64;
  constexpr int NCHK = HALF ? 2 : 4;      // chunks per layer phase
  constexpr int CU = 64 / NCHK;           // hidden units per chunk
  constexpr int BW = HALF ? 32 : 64;      // accumulator block width (output units per lane half)
  constexpr uint32_t LHALF = 16u << 16;   // TMEM address of the second lane half
  constexpr int NL = 2;
  constexpr int NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL, OF = tcw_out_feats(S), CF = tcw_cot_feats(S);
  static_assert(S > 4 && S <= kTcwMaxS, "wide-state tensor-core recurrence: 4 < S <= 10");
  static_assert(L::bytes <= 227 * 1024, "shared memory budget");
  constexpr uint32_t TMEM_COLS = 512;
  // dh_carry ping-pong [layer][parity] | dh_in0 (also d_out . W_out) | direct term [layer] | d z_t
  constexpr uint32_t IN0_COL = 4 * BW, DIR_COL = 5 * BW, DZ_COL = 7 * BW;
  constexpr int SLOT_BYTES = 2 * kATileBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw_tcwb[];
  uint8_t* smem = smem_raw_tcwb + ((1024u - (smem_u32(smem_raw_tcwb) & 1023u)) & 1023u);
  typename L::Bars* bars = reinterpret_cast<typename L::Bars*>(smem + L::OFF_BAR);
  float* wrest = reinterpret_cast<float*>(smem + L::OFF_WREST);
  float* maxb = reinterpret_cast<float*>(smem + L::OFF_MAX);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = (int)p.T;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.wimg);
  const int ew = reinterpret_cast<const int*>(img)[0], eo = reinterpret_cast<const int*>(img)[1], ez = reinterpret_cast<const int*>(img)[2];
  constexpr int KMU = L::KMU, NREST = L::NREST;

  for (int idx = tid; idx < NREST * 64; idx += kBwdThreads) wrest[idx] = p.out_w[(KMU + idx / 64) * 64 + idx % 64];
  if (tid == 0) {
    mbar_init(&bars->full[0], kEpiThreads);
    mbar_init(&bars->full[1], kEpiThreads);
    mbar_init(&bars->empty[0], 1);
    mbar_init(&bars->empty[1], 1);
    mbar_init(&bars->in0, 1);
    mbar_init(&bars->wy, 1);
    mbar_init(&bars->pro, 1);
    mbar_init(&bars->outd, 1);
    mbar_init(&bars->dzr, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const int64_t nitems = (p.B + kTileRows - 1) / kTileRows * SUBS;
  if (tid == 0) {
    mbar_expect_tx(&bars->pro, kWImg + kOutBwdImg + kWzBwdImg);
    bulk_load_1d(smem + L::OFF_W1, img + kImgBwd0 + kWImg, kWImg, &bars->pro);
    bulk_load_1d(smem + L::OFF_WOUT, img + kImgOutBwd, kOutBwdImg, &bars->pro);
    bulk_load_1d(smem + L::OFF_WZ, img + kImgWzBwd, kWzBwdImg, &bars->pro);
  }
  auto load_y = [&](int m) {  // thread 0: Y <- image of W_hh_l0^T (m = 0) / W_hh_l1^T (m = 2); the MMAs reading Y have completed
    mbar_expect_tx(&bars->wy, kWImg);
    bulk_load_1d(smem + L::OFF_Y, img + kImgBwd0 + (size_t)m * kWImg, kWImg, &bars->wy);
  };

  // ---- MMA issue: chunk gc is issued by lane 0 of warp gc % 8 once all 256 threads have written it
  const uint32_t w1 = smem_u32(smem + L::OFF_W1), wy = smem_u32(smem + L::OFF_Y), a0 = smem_u32(smem + L::OFF_A);
  constexpr uint32_t ID64 = idesc_f16(64, MT), ID32 = idesc_f16(32, MT);
  // M = 128, 9 MMAs: acc[128,64] (+)= A_chunk[slots] . W^T[K-groups of chunk c]; wbase = hi tile of the transposed matrix.
  // M = 64, 2 x 18 MMAs: acc_h[64,32] (+)= A_chunk[gate tiles] . W^T[K-groups of chunk c, rows of unit half h]
  auto issue = [&](uint32_t acc, uint32_t slot_base, uint32_t wbase, int c, bool n_is_nh, bool fresh) {
    const uint32_t a_hi = slot_base, a_lo = slot_base + kATileBytes;
    const uint32_t b_hi = wbase, b_lo = wbase + kWTileBytes;
    if (HALF) {
#pragma unroll
      for (uint32_t h = 0; h < 2; ++h) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const int ga = g < 2 ? g : (n_is_nh ? 3 : 2);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t aoff = (uint32_t)(ga >> 1) * 8192u + (uint32_t)(ga & 1) * 64u + (uint32_t)ks * 32u;
            const int gB = c * 6 + g * 2 + ks;
            const uint32_t boff = (uint32_t)(gB >> 2) * 8192u + (uint32_t)(gB & 3) * 32u + h * 4096u;
            const uint64_t dah = umma_desc(a_hi + aoff, 16, 1024, 2), dal = umma_desc(a_lo + aoff, 16, 1024, 2);
            const uint64_t dbh = umma_desc(b_hi + boff, 16, 1024, 2), dbl = umma_desc(b_lo + boff, 16, 1024, 2);
            umma_f16(acc + h * LHALF, dal, dbh, ID32, (fresh && g == 0 && ks == 0) ? 0u : 1u);
            umma_f16(acc + h * LHALF, dah, dbl, ID32, 1u);
            umma_f16(acc + h * LHALF, dah, dbh, ID32, 1u);
          }
        }
      }
      return;
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int aslot = g < 2 ? g : (n_is_nh ? 3 : 2);
      const int gB = c * 3 + g;
      const uint32_t boff = (uint32_t)(gB >> 2) * 8192u + (uint32_t)(gB & 3) * 32u;
      const uint64_t dah = umma_desc(a_hi + aslot * 32, 16, 1024, 2), dal = umma_desc(a_lo + aslot * 32, 16, 1024, 2);
      const uint64_t dbh = umma_desc(b_hi + boff, 16, 1024, 2), dbl = umma_desc(b_lo + boff, 16, 1024, 2);
      umma_f16(acc, dal, dbh, ID64, (fresh && g == 0) ? 0u : 1u);
      umma_f16(acc, dah, dbl, ID64, 1u);
      umma_f16(acc, dah, dbh, ID64, 1u);
    }
  };
  const uint32_t wob = smem_u32(smem + L::OFF_WOUT), wzb = smem_u32(smem + L::OFF_WZ);
  constexpr uint32_t ID16 = idesc_f16(16, MT);
  // d z_t [rows,16] (+)= d_gi_l0 chunk (gates r, u, n) . W_z^T[K-groups of chunk c]; M = 64: into the first lane half only (the
  // second half of a warp takes d z from it by shuffle: 18 fewer MMAs per chunk)
  auto issue_dz = [&](uint32_t slot_base, int c) {
    const uint32_t a_hi = slot_base, a_lo = slot_base + kATileBytes;
    if (HALF) {
#pragma unroll
      for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t aoff = (uint32_t)(g >> 1) * 8192u + (uint32_t)(g & 1) * 64u + (uint32_t)ks * 32u;
          const int gB = c * 6 + g * 2 + ks;
          const uint32_t boff = (uint32_t)(gB >> 2) * 2048u + (uint32_t)(gB & 3) * 32u;
          const uint64_t dah = umma_desc(a_hi + aoff, 16, 1024, 2), dal = umma_desc(a_lo + aoff, 16, 1024, 2);
          const uint64_t dbh = umma_desc(wzb + boff, 16, 1024, 2), dbl = umma_desc(wzb + 3 * 2048 + boff, 16, 1024, 2);
          umma_f16(tmem + DZ_COL, dal, dbh, ID16, (c == 0 && g == 0 && ks == 0) ? 0u : 1u);
          umma_f16(tmem + DZ_COL, dah, dbl, ID16, 1u);
          umma_f16(tmem + DZ_COL, dah, dbh, ID16, 1u);
        }
      }
      return;
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int gB = c * 3 + g;
      const uint32_t boff = (uint32_t)(gB >> 2) * 2048u + (uint32_t)(gB & 3) * 32u;
      const uint64_t dah = umma_desc(a_hi + g * 32, 16, 1024, 2), dal = umma_desc(a_lo + g * 32, 16, 1024, 2);
      const uint64_t dbh = umma_desc(wzb + boff, 16, 1024, 2), dbl = umma_desc(wzb + 3 * 2048 + boff, 16, 1024, 2);
      umma_f16(tmem + DZ_COL, dal, dbh, ID16, (c == 0 && g == 0) ? 0u : 1u);
      umma_f16(tmem + DZ_COL, dah, dbl, ID16, 1u);
      umma_f16(tmem + DZ_COL, dah, dbh, ID16, 1u);
    }
  };
  uint32_t gc = 0;  // chunks produced so far (ring position / phases); uniform over the CTA
  bool pro_ok = false;  // this lane has seen the prologue copies (W_ih_l1^T, W_out tiles) land

  {
    const int quad = warp & 3, cg = warp >> 2;
    const int lh = HALF ? lane >> 4 : 0;              // M = 64: unit half = TMEM lane half of this thread
    const int row = quad * LPQ + (lane & (LPQ - 1));  // row of the CTA's operand tiles / accumulators
    const int t4 = lh * 2 + cg;                       // this thread among the threads of its row
    const bool wr0 = lh == 0;                         // per-row values are computed by every lane half, stored by the first
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    uint8_t* a_ring = smem + L::OFF_A;
    uint32_t ph_in0 = 0, ph_outd = 0, ph_dzr = 0, xb = 0;

    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int64_t tile = item / SUBS;
      const int grow = (int)(item % SUBS) * MT + row;  // row of the 128-row tile of the global layouts
      const int64_t b_raw = tile * kTileRows + grow;
      const bool ok = wr0 && b_raw < p.B;
      const int64_t b = b_raw < p.B ? b_raw : p.B - 1;
      const float* st_tile = p.stash + tile * T * (int64_t)(NL * kStashSlots * 64 * kTileRows) + grow;
      float* dg_tile = p.dg + tile * T * (int64_t)(NL * kDgSlots * 64 * kTileRows) + grow;
      const float* ct_tile = p.ctile + tile * T * (int64_t)(CF * kTileRows) + grow;  // pad rows hold zeros
      const float* ot_tile = p.otile + tile * T * (int64_t)(OF * kTileRows) + grow;
      float* do_tile = p.dout + tile * T * (int64_t)(NOUT * kTileRows) + grow;

      float pv[2][5][8];
      auto load_chunk = [&](float (&dst)[5][8], int tt, int kk, int cc) {
        const float* sk = st_tile + ((int64_t)tt * NL + kk) * (kStashSlots * 64 * kTileRows) +
                          (HALF ? cc * 32 + lh * 16 + cg * 8 : cc * 16 + cg * 8) * kTileRows;
        const float* hk = sk - (int64_t)NL * (kStashSlots * 64 * kTileRows) + kStashH * 64 * kTileRows;  // step tt - 1
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          dst[0][q] = sk[(kStashR * 64 + q) * kTileRows];
          dst[1][q] = sk[(kStashU * 64 + q) * kTileRows];
          dst[2][q] = sk[(kStashN * 64 + q) * kTileRows];
          dst[3][q] = sk[(kStashNhh * 64 + q) * kTileRows];
          dst[4][q] = tt > 0 ? hk[q * kTileRows] : 0.f;
        }
      };
      auto load_ahead = [&](float (&dst)[5][8], int tt, int kk, int cc, int ahead) {
        int lin = ((T - 1 - tt) * NL + (NL - 1 - kk)) * NCHK + cc + ahead;
        const int t2 = T - 1 - lin / (NCHK * NL), k2 = NL - 1 - (lin / NCHK) % NL, c2 = lin % NCHK;
        if (t2 >= 0) load_chunk(dst, t2, k2, c2);
      };
      load_ahead(pv[0], T - 1, NL - 1, 0, 0);
      load_ahead(pv[1], T - 1, NL - 1, 0, 1);

      float dz[S];
#pragma unroll
      for (int s = 0; s < S; ++s) dz[s] = 0.f;
      float sc_prev[NL], scz_prev = 0.f;  // 2^-(row exponent + weight exponent) of the products issued by the previous step
#pragma unroll
      for (int k = 0; k < NL; ++k) sc_prev[k] = 0.f;

      for (int t = T - 1; t >= 0; --t) {
        const bool first = t == T - 1;
        const uint32_t rpar = (uint32_t)(t & 1);
        const float* ct = ct_tile + (int64_t)t * (CF * kTileRows);
        const float* otr = ot_tile + (int64_t)t * (OF * kTileRows);
        TCWB_TRACE(0);
        // every global value this step's d_out phase needs is requested FIRST (one coalesced line each), ahead of the waits
        // for the previous step's MMAs: gP[t+1], eps_t, and this thread's half of the K = 64 operand (gL / gM entries, raw
        // diagonal entries for the floor rule)
        float gpv[S], ev[S], dv[32], rdv[S], rest_gm[NREST > 0 ? NREST : 1];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          gpv[s] = ct[s * kTileRows];
          ev[s] = ct[(2 * S + S * S + s) * kTileRows];
          rdv[s] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) dv[e] = 0.f;
        auto load_half = [&](auto cgh_tag) {
          constexpr int K0 = decltype(cgh_tag)::value * 32;
#pragma unroll
          for (int s = 0; s < S; ++s) {
#pragma unroll
            for (int j = 0; j <= s; ++j) {
              const int k = s * (s + 1) / 2 + j;
              if (k >= K0 && k < K0 + 32) {
                dv[k - K0] = ct[(2 * S + s * S + j) * kTileRows];
                if (j == s) rdv[s] = otr[(S + k) * kTileRows];  // raw (unfloored) diagonal entry
              }
            }
            if (NTRIL + s >= K0 && NTRIL + s < K0 + 32 && s < KMU) dv[NTRIL + s - K0] = ct[(S + s) * kTileRows];  // gM
          }
        };
        if (cg == 0) load_half(std::integral_constant<int, 0>{}); else load_half(std::integral_constant<int, 1>{});
#pragma unroll
        for (int r = 0; r < NREST; ++r) rest_gm[r] = ct[(S + KMU + r) * kTileRows];
        const uint32_t gc_top = gc;  // gc_top - 1 = the last layer-0 chunk of step t + 1
        // ---- cotangent of z_{t+1}: d z through the state columns of W_ih_l0 was accumulated on the tensor pipe by the layer-0
        // chunks of step t + 1 (16 TMEM columns).  Each chunk issues its d z MMAs AHEAD of its carried-dh product and the last one
        // commits them to `dzr`: this wait does not include the 36 - 54 MMAs of the carried product behind them
        if (!first) {
          mbar_wait(&bars->dzr, ph_dzr);
          ph_dzr ^= 1;
          tc_fence_after();
          uint32_t zv[16];
          tmem_ld16_nowait(tl + DZ_COL, zv);
          tmem_ld_wait();
#pragma unroll
          for (int s = 0; s < S; ++s) {
            if (HALF) zv[s] = __shfl_sync(0xffffffffu, zv[s], lane & 15);  // the row's value lives on the first lane half
            dz[s] = fmaf(scz_prev, __uint_as_float(zv[s]), dz[s]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) dz[s] += gpv[s];
        TCWB_TRACE(1);
        // ---------- cotangent of the output projection (kernels/backward.py:300-334) as an MMA operand ----------
        // d_out has n_out = S + S(S+1)/2 entries per row.  64 of them -- every Cholesky entry and the first KMU mu components --
        // form one K = 64 A operand: this thread computes the 32 entries k = 32 cg + e of its row, the two threads of the row
        // agree on a power-of-two row scale (like the d_pre chunks), the entries go into a ring slot as fp16 hi / lo and
        // dh_top (+)= d_out . W_out is 12 MMAs into the dh_in0 columns (dead until the layer-1 chunks of this step write
        // them).  The remaining mu components (one at S = 10) are contracted in FP32 in pass 1.
        float sc_out = 0.f, rest_d[NREST > 0 ? NREST : 1];
        {
          float* dor = do_tile + (int64_t)t * (NOUT * kTileRows);
          auto half = [&](auto cgh_tag) {
            constexpr int K0 = decltype(cgh_tag)::value * 32;
#pragma unroll
            for (int s = 0; s < S; ++s) {
#pragma unroll
              for (int j = 0; j <= s; ++j) {
                const int k = s * (s + 1) / 2 + j;
                if (k >= K0 && k < K0 + 32) {
                  float d = fmaf(dz[s] * ev[j], p.sqrt_dt, dv[k - K0]);
                  if (j == s) d = (rdv[s] >= VISDE_DIAG_MIN || d < 0.f) ? d : 0.f;  // primitives/bounds.py:20
                  dv[k - K0] = d;
                  if (wr0) dor[(S + k) * kTileRows] = d;
                }
              }
              if (NTRIL + s >= K0 && NTRIL + s < K0 + 32 && s < KMU) {
                const float d = fmaf(dz[s], p.dt, dv[NTRIL + s - K0]);
                dv[NTRIL + s - K0] = d;
                if (wr0) dor[s * kTileRows] = d;
              }
            }
          };
          if (cg == 0) half(std::integral_constant<int, 0>{}); else half(std::integral_constant<int, 1>{});
#pragma unroll
          for (int r = 0; r < NREST; ++r) {
            rest_d[r] = fmaf(dz[KMU + r], p.dt, rest_gm[r]);
            if (cg == 0 && wr0) dor[(KMU + r) * kTileRows] = rest_d[r];
          }
          float mxd = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) mxd = fmaxf(mxd, fabsf(dv[e]));
          if (wr0) maxb[(xb * 2 + cg) * 128 + row] = mxd;  // the other lane half computed the same value
          named_bar_sync(1 + quad, 64);
          mxd = fmaxf(mxd, maxb[(xb * 2 + (cg ^ 1)) * 128 + row]);
          xb ^= 1;
          TCWB_TRACE(10);
          const int ed = row_exp_w(mxd);
          const float rsd = exp2i(ed);
          sc_out = exp2i(-(ed + eo));
          const uint32_t slot = gc & 1;
          if (gc >= 2) mbar_wait(&bars->empty[slot], ((gc >> 1) - 1) & 1);
          uint8_t* ahi = a_ring + slot * SLOT_BYTES;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = dv[c * 8 + q] * rsd;
            uint4 hi, lo;
            split8(x, hi, lo);
            if (wr0) {
              *reinterpret_cast<uint4*>(ahi + sw128(row, 4 * cg + c)) = hi;
              *reinterpret_cast<uint4*>(ahi + kATileBytes + sw128(row, 4 * cg + c)) = lo;
            }
          }
          fence_proxy_async();
          tc_fence_before();
          TCWB_TRACE(11);
          mbar_arrive(&bars->full[slot]);
          // ---- Y <- W_hh_l1^T for this step's layer-1 phase (its carried products are issued for t >= 1 only): every MMA of step
          // t + 1 has completed once its last chunk's commit has (in-order tensor pipe) -- long ago by now, so this does not stall
          if (tid == 0 && t >= 1) {
            if (gc_top >= 1) mbar_wait(&bars->empty[(gc_top - 1) & 1], ((gc_top - 1) >> 1) & 1);
            load_y(2);
          }
          if (warp == (int)(gc & 7)) {
            mbar_wait(&bars->full[slot], (gc >> 1) & 1);
            tc_fence_after();
            if (!pro_ok) {  // first use of a prologue-copied tile by this warp
              mbar_wait(&bars->pro, 0);
              tc_fence_after();
              pro_ok = true;
            }
            if (elect_one_sync()) {
              const uint32_t sb = a0 + slot * SLOT_BYTES;
#pragma unroll
              for (uint32_t h = 0; h < (HALF ? 2u : 1u); ++h) {  // M = 64: W_out rows (hidden units) of half h, N = 32
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint64_t dah = umma_desc(sb + j * 32, 16, 1024, 2), dal = umma_desc(sb + kATileBytes + j * 32, 16, 1024, 2);
                  const uint64_t dbh = umma_desc(wob + h * 4096u + j * 32, 16, 1024, 2);
                  const uint64_t dbl = umma_desc(wob + 64 * 128 + h * 4096u + j * 32, 16, 1024, 2);
                  umma_f16(tmem + h * LHALF + IN0_COL, dal, dbh, HALF ? ID32 : ID64, j > 0 ? 1u : 0u);
                  umma_f16(tmem + h * LHALF + IN0_COL, dah, dbl, HALF ? ID32 : ID64, 1u);
                  umma_f16(tmem + h * LHALF + IN0_COL, dah, dbh, HALF ? ID32 : ID64, 1u);
                }
              }
              umma_commit(&bars->empty[slot]);
              umma_commit(&bars->outd);
            }
            __syncwarp();
          }
          ++gc;
        }
        float sc_in = 0.f;

#pragma unroll
        for (int k = NL - 1; k >= 0; --k) {
          if (k == 0) {
            TCWB_TRACE(5);
            mbar_wait(&bars->in0, ph_in0);
            ph_in0 ^= 1;
            tc_fence_after();
            TCWB_TRACE(6);
          }
          // ---------- pass 1: dh of this thread's 8 NCHK units, row maximum ----------
          float2 dh2[NCHK * 4];  // dh of the q-th unit of chunk c lives in dh2[c * 4 + q / 2].{x, y}
#pragma unroll
          for (int q = 0; q < NCHK * 4; ++q) dh2[q] = make_float2(0.f, 0.f);
          if (k == 1) {
            // d_out . W_out has landed in the dh_in0 columns; the mu components outside the MMA operand are contracted here
            mbar_wait(&bars->outd, ph_outd);
            ph_outd ^= 1;
            tc_fence_after();
#pragma unroll
            for (int r = 0; r < NREST; ++r) {
              const float* wr = wrest + r * 64 + (HALF ? lh * 16 + cg * 8 : cg * 8);
              const float2 d2 = make_float2(rest_d[r], rest_d[r]);
#pragma unroll
              for (int c = 0; c < NCHK; ++c) {
                const float4 wa = *reinterpret_cast<const float4*>(wr + c * CU);
                const float4 wb = *reinterpret_cast<const float4*>(wr + c * CU + 4);
                fma2(dh2[c * 4 + 0], make_float2(wa.x, wa.y), d2);
                fma2(dh2[c * 4 + 1], make_float2(wa.z, wa.w), d2);
                fma2(dh2[c * 4 + 2], make_float2(wb.x, wb.y), d2);
                fma2(dh2[c * 4 + 3], make_float2(wb.z, wb.w), d2);
              }
            }
          }
          if (k == 1) TCWB_TRACE(2); else TCWB_TRACE(7);
          float mx = 0.f;
#pragma unroll
          for (int c = 0; c < NCHK; ++c) {
            const int j0 = c * 16 + cg * 8;  // column of this thread's units of chunk c inside an accumulator block
            uint32_t va[8], vd[8], vi[8];
            if (!first) {
              tmem_ld8_nowait(tl + (uint32_t)(k * 2 + rpar) * BW + j0, va);
              tmem_ld8_nowait(tl + DIR_COL + (uint32_t)k * BW + j0, vd);
            }
            // k = 0: dh_in0 of this step; k = 1: d_out . W_out of this step (same columns, see the d_out phase)
            tmem_ld8_nowait(tl + IN0_COL + j0, vi);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float& dref = (q & 1) ? dh2[c * 4 + q / 2].y : dh2[c * 4 + q / 2].x;
              float v = dref;
              if (!first) v += fmaf(sc_prev[k], __uint_as_float(va[q]), __uint_as_float(vd[q]));
              v = fmaf(k == 0 ? sc_in : sc_out, __uint_as_float(vi[q]), v);
              dref = v;
              mx = fmaxf(mx, fabsf(v));
            }
          }
          // ---------- the threads of the row agree on the power-of-two scale ----------
          if (HALF) {
            maxb[(xb * 4 + t4) * 64 + row] = mx;
            named_bar_sync(1 + quad, 64);
            mx = fmaxf(fmaxf(mx, maxb[(xb * 4 + (t4 ^ 1)) * 64 + row]),
                       fmaxf(maxb[(xb * 4 + (t4 ^ 2)) * 64 + row], maxb[(xb * 4 + (t4 ^ 3)) * 64 + row]));
          } else {
            maxb[(xb * 2 + cg) * 128 + row] = mx;
            named_bar_sync(1 + quad, 64);
            mx = fmaxf(mx, maxb[(xb * 2 + (cg ^ 1)) * 128 + row]);
          }
          xb ^= 1;
          if (k == 1) TCWB_TRACE(3); else TCWB_TRACE(8);
          const int er = row_exp_w(mx);
          const float rs = exp2i(er);
          const float sc_this = exp2i(-(er + ew));

          // ---------- pass 2: gate cotangents, dg, direct term, A-operand chunks ----------
          float* dg_k = dg_tile + ((int64_t)t * NL + k) * (kDgSlots * 64 * kTileRows);
#pragma unroll
          for (int c = 0; c < NCHK; ++c, ++gc) {
            const int j0 = c * 16 + cg * 8;                                      // accumulator column
            const int ju = HALF ? c * 32 + lh * 16 + cg * 8 : c * 16 + cg * 8;  // hidden unit
            float cr[8], cu[8], cn[8], cnh[8], chp[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              cr[q] = pv[c & 1][0][q]; cu[q] = pv[c & 1][1][q]; cn[q] = pv[c & 1][2][q];
              cnh[q] = pv[c & 1][3][q]; chp[q] = pv[c & 1][4][q];
            }
            load_ahead(pv[c & 1], t, k, c, 2);
            float dr_[8], du_[8], dn_[8], dnh_[8];
            uint32_t dirv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int i = ju + q;
              const float r = cr[q], u = cu[q], n = cn[q], nhh = cnh[q], hp = chp[q];
              const float dhv = (q & 1) ? dh2[c * 4 + q / 2].y : dh2[c * 4 + q / 2].x;
              const float dnp = dhv * (1.f - u) * (1.f - n * n);
              const float dup = dhv * (hp - n) * u * (1.f - u);
              const float drp = dnp * nhh * r * (1.f - r);
              const float dnh = dnp * r;
              dirv[q] = __float_as_uint(dhv * u);
              dg_k[(0 * 64 + i) * kTileRows] = drp;
              dg_k[(1 * 64 + i) * kTileRows] = dup;
              dg_k[(2 * 64 + i) * kTileRows] = dnp;
              dg_k[(3 * 64 + i) * kTileRows] = dnh;
              dr_[q] = drp * rs; du_[q] = dup * rs; dn_[q] = dnp * rs; dnh_[q] = dnh * rs;
            }
            tmem_st8(tl + DIR_COL + (uint32_t)k * BW + j0, dirv);
            if (k == 0 && c == 0 && tid == 0 && t >= 1) {
              // Y <- W_hh_l0^T once the carried products of the layer-1 phase (issued BEHIND the dh_in0 products that `in0` covers)
              // have finished reading W_hh_l1^T: about a microsecond after `in0`, i.e. by now, so this wait does not stall
              mbar_wait(&bars->empty[(gc - 1) & 1], ((gc - 1) >> 1) & 1);
              load_y(0);
            }
            const uint32_t slot = gc & 1;
            if (gc >= 2) mbar_wait(&bars->empty[slot], ((gc >> 1) - 1) & 1);
            uint8_t* ahi = a_ring + slot * SLOT_BYTES;
            {
              // M = 128: one K = 64 tile, 16-byte group 2 g + cg of gate g.  M = 64: two K = 64 tiles of 64 rows (gates r | u, then
              // n | n_hh), group 4 (g & 1) + t4 of tile g >> 1
              const uint32_t o0 = HALF ? sw128(row, t4) : sw128(row, 0 + cg);
              const uint32_t o1 = HALF ? sw128(row, 4 + t4) : sw128(row, 2 + cg);
              const uint32_t o2 = HALF ? 8192u + sw128(row, t4) : sw128(row, 4 + cg);
              const uint32_t o3 = HALF ? 8192u + sw128(row, 4 + t4) : sw128(row, 6 + cg);
              uint4 hi, lo;
              split8(dr_, hi, lo);
              *reinterpret_cast<uint4*>(ahi + o0) = hi;
              *reinterpret_cast<uint4*>(ahi + kATileBytes + o0) = lo;
              split8(du_, hi, lo);
              *reinterpret_cast<uint4*>(ahi + o1) = hi;
              *reinterpret_cast<uint4*>(ahi + kATileBytes + o1) = lo;
              split8(dn_, hi, lo);
              *reinterpret_cast<uint4*>(ahi + o2) = hi;
              *reinterpret_cast<uint4*>(ahi + kATileBytes + o2) = lo;
              split8(dnh_, hi, lo);
              *reinterpret_cast<uint4*>(ahi + o3) = hi;
              *reinterpret_cast<uint4*>(ahi + kATileBytes + o3) = lo;
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->full[slot]);
            if (warp == (int)(gc & 7)) {
              mbar_wait(&bars->full[slot], (gc >> 1) & 1);
              if (!pro_ok) {  // first use of a prologue-copied tile (W_ih_l1^T, W_z) by this warp
                mbar_wait(&bars->pro, 0);
                pro_ok = true;
              }
              tc_fence_after();
              const uint32_t sb = a0 + slot * SLOT_BYTES;
              // the product the NEXT phase waits for goes first and gets its own commit: dh_in0 (layer 1 -> `in0`), d z_t (layer 0
              // -> `dzr`); the carried dh, read one step later, follows
              if (elect_one_sync()) {
                if (k == 1) issue(tmem + IN0_COL, sb, w1, c, false, c == 0);
                else issue_dz(sb, c);
                if (c == NCHK - 1) umma_commit(k == 1 ? &bars->in0 : &bars->dzr);
              }
              __syncwarp();
              if (t > 0) {
                mbar_wait(&bars->wy, k == 1 ? 0u : 1u);  // Y holds W_hh_l1^T in the layer-1 phase (even completions of wy),
                tc_fence_after();                        // W_hh_l0^T in the layer-0 phase (odd)
              }
              if (elect_one_sync()) {
                if (t > 0) issue(tmem + (uint32_t)(k * 2 + ((t & 1) ^ 1)) * BW, sb, wy, c, true, c == 0);
                umma_commit(&bars->empty[slot]);
              }
              __syncwarp();
            }
          }
          tmem_st_wait();
          if (k == 1) TCWB_TRACE(4); else TCWB_TRACE(9);
          sc_prev[k] = sc_this;
          if (k == 1) sc_in = sc_this;
          else scz_prev = exp2i(-(er + ez));
        }
      }
      // grad_x0 = d z_0 + g_paths[:, 0]; the state-column part of d z_0 sits in TMEM behind the last chunk's MMAs
      {
        mbar_wait(&bars->dzr, ph_dzr);
        ph_dzr ^= 1;
        mbar_wait(&bars->empty[(gc - 1) & 1], ((gc - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t zv[16];
        tmem_ld16_nowait(tl + DZ_COL, zv);
        tmem_ld_wait();
        tc_fence_before();
        if (ok && cg == 0) {
#pragma unroll
          for (int s = 0; s < S; ++s)
            p.grad_x0[b * S + s] = fmaf(scz_prev, __uint_as_float(zv[s]), dz[s]) + p.g_paths[b * (T + 1) * S + s];
        }
      }
    }
    if (gc >= 2) mbar_wait(&bars->empty[(gc - 2) & 1], ((gc - 2) >> 1) & 1);
    if (gc >= 1) mbar_wait(&bars->empty[(gc - 1) & 1], ((gc - 1) >> 1) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int S>
int launch_bwd_tcw(const PathParams& p, cudaStream_t st) {
  const size_t smem = TcwBwdSmem<S>::bytes;
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_tcw_kernel<S, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_tcw_kernel<S, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (p.B + kTileRows - 1) / kTileRows;
  if (tcw_half_tiles(ntiles, sms))
    path_bwd_tcw_kernel<S, 64><<<(unsigned)(2 * ntiles), kBwdThreads, smem, st>>>(p);
  else
    path_bwd_tcw_kernel<S, 128><<<(unsigned)(ntiles < sms ? ntiles : sms), kBwdThreads, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// thin gradient pieces of the wide family from the tiled buffers (thread = trajectory row, lanes = rows: every access a
// full 128-byte line), per-tile partial records, fixed-order sum over tiles.
//   part A: bias sums (all dg features), sum_t d_gi_l0 (theta columns), dW_ih_l0[:, :S] = sum d_gi (x) z_t
//   part B: dW_out = sum d_out (x) h_top, db_out = sum d_out
// record of one tile: [F] bias sums | [192][S] dW_z | [NOUT][64] dW_out | [NOUT] db_out
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int tcw_part_floats(int NL, int S) {
  const int nout = S + S * (S + 1) / 2;
  return NL * kDgSlots * 64 + 192 * S + nout * 64 + nout;
}
constexpr int kTwFeat = 4;    // dg features per thread in part A

// volatile: the compiler must not sink these loads behind the FMAs of an earlier step (it re-serialises an unrolled batch otherwise)
__device__ __forceinline__ float ldg_batch(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

template <int S>
__global__ void __launch_bounds__(256) tcw_thin_a_kernel(const float* __restrict__ dg, const float* __restrict__ otile,
                                                         const float* __restrict__ paths, int64_t B, int T, float* __restrict__ sdg,
                                                         float* __restrict__ part) {
  constexpr int NL = 2, F = NL * kDgSlots * 64, OF = tcw_out_feats(S), NTRIL = S * (S + 1) / 2;
  const int64_t tb = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, rq = w & 3, par = w >> 2;
  const int row = rq * 32 + lane;
  const int64_t b_raw = tb * kTileRows + row;
  const bool ok = b_raw < B;
  const int64_t b = ok ? b_raw : B - 1;
  float* prec = part + tb * tcw_part_floats(NL, S);
  __shared__ float red[2][4][kTwFeat * (1 + S)];
  const int f0 = (blockIdx.y * 2 + par) * kTwFeat;
  const int64_t tstride = (int64_t)F * kTileRows;
  const float* src = dg + tb * T * tstride + (int64_t)f0 * kTileRows + row;
  const bool wz = f0 < 192;  // layer-0 slots r, u, n feed the state columns of W_ih_l0
  // z_t: paths[:, 0] for t = 0, the z_{t+1} entry of step record t - 1 afterwards
  const float* zrec = otile + tb * T * (int64_t)(OF * kTileRows) + (int64_t)(S + NTRIL) * kTileRows + row;
  float acc[kTwFeat], accz[kTwFeat][S];
#pragma unroll
  for (int j = 0; j < kTwFeat; ++j) {
    acc[j] = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s) accz[j][s] = 0.f;
  }
  // four grid steps per trip, all their loads requested before the first use (the pass is HBM-latency bound: 74 % long_scoreboard
  // stalls with one step in flight)
  constexpr int TB = 4;
  for (int t0 = 0; t0 < T; t0 += TB) {
    float v[TB][kTwFeat], z[TB][S];
#pragma unroll
    for (int u = 0; u < TB; ++u) {
      const int t = t0 + u;
      const bool in = t < T;
      const float* st = src + (in ? t : 0) * tstride;
#pragma unroll
      for (int j = 0; j < kTwFeat; ++j) v[u][j] = ldg_batch(st + j * kTileRows);
      if (wz) {  // block-uniform
        const float* zp = (t >= 1 && in) ? zrec + (int64_t)(t - 1) * (OF * kTileRows) : paths + b * (int64_t)(T + 1) * S;
        const int zs = (t >= 1 && in) ? kTileRows : 1;
#pragma unroll
        for (int s = 0; s < S; ++s) z[u][s] = ldg_batch(zp + s * zs);
      } else {
#pragma unroll
        for (int s = 0; s < S; ++s) z[u][s] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < TB; ++u) {
      const bool in = t0 + u < T;
#pragma unroll
      for (int j = 0; j < kTwFeat; ++j) v[u][j] = in ? v[u][j] : 0.f;
      if (t0 + u == 0 && !ok) {  // pad rows have no z_0 (the clamped row's was read): their d_pre is exactly zero anyway
#pragma unroll
        for (int s = 0; s < S; ++s) z[u][s] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < TB; ++u)
#pragma unroll
      for (int j = 0; j < kTwFeat; ++j) {
        acc[j] += v[u][j];
#pragma unroll
        for (int s = 0; s < S; ++s) accz[j][s] = fmaf(v[u][j], z[u][s], accz[j][s]);
      }
  }
  if (wz && ok) {
#pragma unroll
    for (int j = 0; j < kTwFeat; ++j) sdg[b * 192 + f0 + j] = acc[j];
  }
#pragma unroll
  for (int j = 0; j < kTwFeat; ++j) {
    float a = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[par][rq][j] = a;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float c = accz[j][s];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) red[par][rq][kTwFeat + j * S + s] = c;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * kTwFeat * (1 + S); idx += blockDim.x) {
    const int pp = idx / (kTwFeat * (1 + S)), q = idx % (kTwFeat * (1 + S));
    const float a = (red[pp][0][q] + red[pp][1][q]) + (red[pp][2][q] + red[pp][3][q]);
    const int fb = (blockIdx.y * 2 + pp) * kTwFeat;
    if (q < kTwFeat) prec[fb + q] = a;
    else if (fb < 192) prec[F + (fb + (q - kTwFeat) / S) * S + (q - kTwFeat) % S] = a;
  }
}

// dW_out[m][i] = sum_{rows, t} d_out[m] h_top[i] and db_out[m] = sum d_out[m]: an [80 x 64] x K SGEMM with K = (tile, t, row).
// Both operands are K-contiguous in the tiled buffers ([feature][128 rows] per (tile, t)), so a CTA streams whole
// (tile, t) records through a cp.async double buffer (65 + 64 lines of 512 B each, read exactly once) and every thread keeps a
// 5 x 4 block of the product in registers (80 FFMA per 9 LDS.128); per-CTA partials are summed in a fixed order afterwards.
constexpr int kTbThreads = 256, kTbPitch = 132;  // pitch 132 floats: quarter-warp LDS.128 of 8 consecutive rows hit 32 distinct banks
constexpr int kTbMaxCtas = 148;  // 152 KB of staging per CTA: one CTA per SM

__device__ __forceinline__ void cp_async16_w(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int S>
__global__ void __launch_bounds__(kTbThreads) tcw_thin_b_kernel(const float* __restrict__ dout, const float* __restrict__ stash,
                                                                int64_t nrec, float* __restrict__ cta_part) {
  constexpr int NL = 2, NOUT = S + S * (S + 1) / 2, MP = 80;  // m padded to 16 groups of 5
  extern __shared__ __align__(16) float tb_s[];  // [2 stages][MP + 64][kTbPitch]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t sstride = (int64_t)NL * kStashSlots * 64 * kTileRows;
  const int64_t hoff = (int64_t)((NL - 1) * kStashSlots + kStashH) * 64 * kTileRows;
  constexpr int STAGE = (MP + 64) * kTbPitch;
  // rows NOUT..MP-1 of the d_out stage are never copied: zero them once
  for (int i = tid; i < 2 * STAGE; i += kTbThreads) tb_s[i] = 0.f;
  __syncthreads();
  const int64_t per = (nrec + gridDim.x - 1) / gridDim.x;
  const int64_t r_beg = (int64_t)blockIdx.x * per, r_end = r_beg + per < nrec ? r_beg + per : nrec;
  auto stage = [&](int buf, int64_t rec) {  // record = (tile, t): d_out [NOUT][128] and h_top [64][128], contiguous lines
    float* dst = tb_s + buf * STAGE;
    const float* d = dout + rec * (int64_t)(NOUT * kTileRows);
    const float* h = stash + rec * sstride + hoff;
    for (int c = tid; c < (NOUT + 64) * 32; c += kTbThreads) {
      const int f = c >> 5, q = c & 31;
      const float* src = f < NOUT ? d + f * kTileRows + 4 * q : h + (f - NOUT) * kTileRows + 4 * q;
      cp_async16_w(dst + (f < NOUT ? f : MP + f - NOUT) * kTbPitch + 4 * q, src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float acc[5][4], accd[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    accd[a] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  }
  if (r_beg < r_end) stage(0, r_beg);
  int cur = 0;
  for (int64_t rec = r_beg; rec < r_end; ++rec) {
    if (rec + 1 < r_end) {
      stage(cur ^ 1, rec + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* As = tb_s + cur * STAGE + (ty * 5) * kTbPitch;           // d_out rows ty*5 .. ty*5+4
    const float* Bs = tb_s + cur * STAGE + (MP + tx) * kTbPitch;          // h_top units tx, tx+16, tx+32, tx+48
#pragma unroll 4
    for (int r = 0; r < kTileRows; r += 4) {
      float4 av[5], bv[4];
#pragma unroll
      for (int a = 0; a < 5; ++a) av[a] = *reinterpret_cast<const float4*>(As + a * kTbPitch + r);
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = *reinterpret_cast<const float4*>(Bs + b * 16 * kTbPitch + r);
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        if (tx == 0) accd[a] += (av[a].x + av[a].y) + (av[a].z + av[a].w);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          acc[a][b] = fmaf(av[a].x, bv[b].x, acc[a][b]);
          acc[a][b] = fmaf(av[a].y, bv[b].y, acc[a][b]);
          acc[a][b] = fmaf(av[a].z, bv[b].z, acc[a][b]);
          acc[a][b] = fmaf(av[a].w, bv[b].w, acc[a][b]);
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  // per-CTA record: [NOUT][64] dW_out | [NOUT] db_out
  float* out = cta_part + (int64_t)blockIdx.x * (NOUT * 64 + NOUT);
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    const int m = ty * 5 + a;
    if (m < NOUT) {
#pragma unroll
      for (int b = 0; b < 4; ++b) out[m * 64 + tx + 16 * b] = acc[a][b];
      if (tx == 0) out[NOUT * 64 + m] = accd[a];
    }
  }
}

// fixed-order sum of the per-CTA records of tcw_thin_b_kernel
__global__ void tcw_thin_b_reduce_kernel(const float* __restrict__ cta_part, int nctas, int n, float* __restrict__ out_w,
                                         float* __restrict__ out_b, int n_w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float acc = 0.f;
  for (int c = 0; c < nctas; ++c) acc += cta_part[(int64_t)c * n + idx];
  if (idx < n_w) out_w[idx] = acc;
  else out_b[idx - n_w] = acc;
}

struct TcwReduceArgs {
  const float* part;
  int ntile, NL, S, n_out, ld0;
  float* b_ih[VISDE_MAX_LAYERS];
  float* b_hh[VISDE_MAX_LAYERS];
  float* w_ih0;
  float* out_w;
  float* out_b;
};
__global__ void tcw_thin_reduce_kernel(TcwReduceArgs a) {
  const int total = tcw_part_floats(a.NL, a.S);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float acc = 0.f;
  for (int c = 0; c < a.ntile; ++c) acc += a.part[(int64_t)c * total + idx];
  const int F = a.NL * kDgSlots * 64;
  int off = idx;
  if (off < F) {
    const int k = off / (kDgSlots * 64), slot = (off / 64) % kDgSlots, i = off % 64;
    if (slot < 2) {
      a.b_ih[k][slot * 64 + i] = acc;
      a.b_hh[k][slot * 64 + i] = acc;
    } else if (slot == 2) {
      a.b_ih[k][128 + i] = acc;
    } else {
      a.b_hh[k][128 + i] = acc;
    }
    return;
  }
  off -= F;
  if (off < 192 * a.S) {
    a.w_ih0[(int64_t)(off / a.S) * a.ld0 + off % a.S] = acc;
    return;
  }
  // dW_out / db_out come from tcw_thin_b_reduce_kernel (the per-tile records keep their slots unused)
}

template <int S>
int launch_thin_tcw(const PathParams& p, const visde_weight_grads* gw, float* partials, cudaStream_t st) {
  const int64_t ntile = (p.B + kTileRows - 1) / kTileRows;
  constexpr int F = 2 * kDgSlots * 64, NOUT = S + S * (S + 1) / 2;
  tcw_thin_a_kernel<S><<<dim3((unsigned)ntile, F / (2 * kTwFeat)), 256, 0, st>>>(p.dg, p.otile, p.paths, p.B, (int)p.T, p.sdg, partials);
  VISDE_CUDA_CHECK(cudaGetLastError());
  // dW_out / db_out: per-CTA records behind the per-tile records of part A
  float* cta_part = partials + (size_t)ntile * tcw_part_floats(2, S);
  const int64_t nrec = ntile * p.T;
  const int nctas = (int)(nrec < kTbMaxCtas ? nrec : kTbMaxCtas);
  const size_t smem = sizeof(float) * 2 * (80 + 64) * kTbPitch;
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(tcw_thin_b_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  tcw_thin_b_kernel<S><<<nctas, kTbThreads, smem, st>>>(p.dout, p.stash, nrec, cta_part);
  VISDE_CUDA_CHECK(cudaGetLastError());
  const int n = NOUT * 64 + NOUT;
  tcw_thin_b_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(cta_part, nctas, n, gw->out_w, gw->out_b, NOUT * 64);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace

#ifdef VISDE_TCW_TRACE
extern "C" int visde_debug_tcw_trace_bwd(long long* out) {
  return cudaMemcpyFromSymbol(out, g_tcw_trace_bwd, sizeof(g_tcw_trace_bwd)) == cudaSuccess ? 0 : -2;
}
#endif

size_t tcw_image_bytes() { return (kImgBytes + 255) / 256 * 256; }

size_t tcw_thin_partial_floats(int64_t B, int NL, int S) {
  const int nout = S + S * (S + 1) / 2;
  return (size_t)((B + kTileRows - 1) / kTileRows) * tcw_part_floats(NL, S) + (size_t)kTbMaxCtas * (nout * 64 + nout);
}

// p.stash / p.dg / p.dout / p.otile / p.ctile are tiled buffers; p.wimg holds the backward images
int launch_path_bwd_tcw(const PathParams& p, cudaStream_t st) {
  // cotangent record [tile][t][3S + S*S][128]: gP[t+1] | gM | gL | eps
  const int S = p.S, CF = tcw_cot_feats(S);
  // two launches: the three S-wide sources share one (16 grid steps per block: 640-byte row segments), gL has its own (S*S wide)
  const float* src[3] = {p.g_paths + S, p.g_means, p.eps};
  const int64_t bs[3] = {(p.T + 1) * (int64_t)S, p.T * (int64_t)S, p.T * (int64_t)S};
  const int64_t ts[3] = {S, S, S};
  const int F[3] = {S, S, S}, fo[3] = {0, S, 2 * S + S * S};
  int rc = launch_tcw_tile_multi(src, bs, ts, F, fo, 3, p.B, p.T, p.ctile, CF, st);
  if (rc) return rc;
  if ((rc = launch_tcw_tile(p.g_chol, p.B, p.T, S * S, p.T * (int64_t)S * S, S * S, p.ctile, CF, 2 * S, st))) return rc;
  switch (S) {
    case 5: return launch_bwd_tcw<5>(p, st);
    case 6: return launch_bwd_tcw<6>(p, st);
    case 7: return launch_bwd_tcw<7>(p, st);
    case 8: return launch_bwd_tcw<8>(p, st);
    case 9: return launch_bwd_tcw<9>(p, st);
    case 10: return launch_bwd_tcw<10>(p, st);
  }
  set_error("wide-state tensor-core recurrence: unsupported state dim %d", S);
  return VISDE_EINVAL;
}

int launch_tcw_thin_grads(const PathParams& p, const visde_weight_grads* gw, float* partials, size_t partial_floats,
                          cudaStream_t st) {
  if (tcw_thin_partial_floats(p.B, p.NL, p.S) > partial_floats) {
    set_error("tcw thin gradients: workspace too small");
    return VISDE_EWORKSPACE;
  }
  int rc;
  switch (p.S) {
    case 5: rc = launch_thin_tcw<5>(p, gw, partials, st); break;
    case 6: rc = launch_thin_tcw<6>(p, gw, partials, st); break;
    case 7: rc = launch_thin_tcw<7>(p, gw, partials, st); break;
    case 8: rc = launch_thin_tcw<8>(p, gw, partials, st); break;
    case 9: rc = launch_thin_tcw<9>(p, gw, partials, st); break;
    case 10: rc = launch_thin_tcw<10>(p, gw, partials, st); break;
    default: set_error("tcw thin gradients: unsupported state dim %d", p.S); return VISDE_EINVAL;
  }
  if (rc) return rc;
  TcwReduceArgs a{};
  a.part = partials;
  a.ntile = (int)((p.B + kTileRows - 1) / kTileRows);
  a.NL = p.NL;
  a.S = p.S;
  a.n_out = p.n_out;
  a.ld0 = p.S + p.C + p.P;
  for (int k = 0; k < p.NL; ++k) {
    a.b_ih[k] = gw->b_ih[k];
    a.b_hh[k] = gw->b_hh[k];
  }
  a.w_ih0 = gw->w_ih[0];
  a.out_w = gw->out_w;
  a.out_b = gw->out_b;
  const int total = tcw_part_floats(p.NL, p.S);
  tcw_thin_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace visde
