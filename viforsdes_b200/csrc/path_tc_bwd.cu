// Tensor-core BPTT for large batches (B >> #SM, H = 64): K2c path_bwd_tc.
//
// Mirror of path_tc.cu in reverse time.  A CTA owns 128 trajectories (MMA M = TMEM lanes); per layer
// and step the gate cotangents d_pre = (d r_pre, d u_pre, d n_pre, d n_hh) [128, 4 x 64] are the A
// operand of
//     dhc_k  [128,64] = d_gh_k . W_hh_k      (carried to step t-1, ping-pong accumulators)
//     dh_in0 [128,64] = d_gi_1 . W_ih_1      (handed to layer 0 of the same step)
// issued as tcgen05.mma kind::f16 (M = 128, N = 64, K = 16 per instruction) from shared-memory tiles,
// with the same fp16 hi/lo 3-pass split as the forward.  Gradients have no fixed range, so every
// trajectory row is scaled by its own power of two per layer-step, 2^e with max|dh| 2^e in
// [2^9, 2^10): the two epilogue threads of a row agree on it through a 64-thread named barrier, and
// the accumulator is read back with 2^-(e + a_w).  d_pre is produced in chunks of 16 hidden units
// (a 2-slot ring of [128 x 64] fp16 hi/lo tiles: 4 K-groups = the 4 slots of the chunk's units), so
// the MMAs of chunk c run while the epilogue computes chunk c+1 and the A operand never needs more
// than 64 KB next to the 144 KB of resident transposed weights.  The direct term dh * u of the GRU
// cell is parked in TMEM next to the accumulators (tcgen05.st) instead of registers.
//
// Inputs / outputs use the row-fastest tiled layouts of the forward: stash [tile][t][NL][5][H][128]
// is read with one coalesced line per warp access, d_pre is written as dg [tile][t][NL][4][H][128].
// The reductions over (b, t) (biases, state columns of W_ih_l0, W_out, sum_t d_gi for theta) are
// left to the time-parallel kernels that read dg.
#include "path_tc.cuh"

namespace visde {
namespace {

constexpr int kBwdThreads = 256;  // 8 warps = 255 registers per thread; the warps take turns as MMA issuer
constexpr int kRowExp = 9;  // rows are scaled so that max|dh| 2^e is in [2^9, 2^10)

template <int NL, int S>
struct TcBwdSmem {
  static constexpr int NMAT = 2 * NL - 1;  // W_hh_l0^T, W_ih_l1^T, W_hh_l1^T
  static constexpr int CZ = (3 * S <= 4) ? 4 : (3 * S <= 8 ? 8 : 12);
  static constexpr int OFF_W = 0;                               // [NMAT][hi, lo][3 K-blocks][64][128 B]
  static constexpr int OFF_A = OFF_W + NMAT * 2 * kWTileBytes;  // ring [2][hi, lo][128][128 B]
  static constexpr int OFF_WOUT = OFF_A + 4 * kATileBytes;      // float [64][16]: W_out[m][i] at [i][m]
  static constexpr int OFF_WZ = OFF_WOUT + 64 * 16 * 4;         // float [64][CZ]: W_ih_l0[g*64+i][s] at [i][g*S+s]
  static constexpr int OFF_MAX = OFF_WZ + 64 * CZ * 4;          // float [2 buffers][2 cg][128]
  static constexpr int OFF_DZX = OFF_MAX + 2 * 2 * 128 * 4;     // float [2 cg][128][S]
  static constexpr int OFF_BAR = OFF_DZX + 2 * 128 * S * 4;
  struct Bars {
    uint64_t full[2], empty[2], in0;
    uint32_t tmem_base;
    uint32_t amax;
  };
  static constexpr size_t bytes = OFF_BAR + sizeof(Bars) + 1024;
};

__device__ __forceinline__ int row_exp(float mx) {
  const uint32_t bits = __float_as_uint(mx);
  if (bits == 0u) return 0;
  int e = kRowExp - ((int)(bits >> 23) - 127);
  e = e > 100 ? 100 : e;
  return e < -100 ? -100 : e;
}

template <int NL, int S>
__global__ void __launch_bounds__(kBwdThreads, 1) path_bwd_tc_kernel(PathParams p) {
  using L = TcBwdSmem<NL, S>;
  constexpr int NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL, CZ = L::CZ, NMAT = L::NMAT;
  static_assert(NOUT <= 16, "W_out columns are staged as 16 floats per unit");
  constexpr uint32_t TMEM_COLS = NL == 2 ? 512 : 256;
  // accumulators: dhc_k ping-pong [k][parity] (64 columns each), dh_in0; then the parked direct terms
  constexpr uint32_t IN0_COL = NL == 2 ? 256 : 0, DIR_COL = NL == 2 ? 320 : 128;
  constexpr int SLOT_BYTES = 2 * kATileBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw_tcb[];
  uint8_t* smem = smem_raw_tcb + ((1024u - (smem_u32(smem_raw_tcb) & 1023u)) & 1023u);
  typename L::Bars* bars = reinterpret_cast<typename L::Bars*>(smem + L::OFF_BAR);
  float* woutc = reinterpret_cast<float*>(smem + L::OFF_WOUT);
  float* wzc = reinterpret_cast<float*>(smem + L::OFF_WZ);
  float* maxb = reinterpret_cast<float*>(smem + L::OFF_MAX);
  float* dzx = reinterpret_cast<float*>(smem + L::OFF_DZX);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld0 = S + p.C + p.P;
  const int T = (int)p.T;

  // ---- weight scaling exponent (shared by the three matrices: their products meet in dh)
  if (tid == 0) bars->amax = 0u;
  __syncthreads();
  {
    float mx = 0.f;
    for (int m = 0; m < NMAT; ++m) {
      const float* src = m == 0 ? p.w_hh[0] : (m == 1 ? p.w_ih[1] : p.w_hh[1]);
      for (int idx = tid; idx < 192 * 64; idx += kBwdThreads) mx = fmaxf(mx, fabsf(src[idx]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) atomicMax(&bars->amax, __float_as_uint(mx));
  }
  __syncthreads();
  const int ew = scale_exp(bars->amax);
  const float w_scale = exp2i(ew);

  // ---- resident transposed weight tiles.  B operand rows = input / hidden index i (N = 64); K runs over
  // (chunk c of 16 units, gate g, unit in chunk): K-group gB = c * 3 + g sits in K-block gB / 4 at byte
  // column (gB % 4) * 32 of the 128-byte swizzled row
  for (int m = 0; m < NMAT; ++m) {
    const float* src = m == 0 ? p.w_hh[0] : (m == 1 ? p.w_ih[1] : p.w_hh[1]);
    uint8_t* thi = smem + L::OFF_W + (m * 2) * kWTileBytes;
    uint8_t* tlo = thi + kWTileBytes;
    for (int idx = tid; idx < 64 * 24; idx += kBwdThreads) {
      const int i = idx & 63, hg = idx >> 6;  // hg = gB * 2 + half
      const int gB = hg >> 1, half = hg & 1, c = gB / 3, g = gB % 3;
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = src[(g * 64 + c * 16 + half * 8 + q) * 64 + i] * w_scale;
      uint4 hi, lo;
      split8(x, hi, lo);
      const uint32_t off = (uint32_t)(gB >> 2) * 8192u + sw128(i, (gB & 3) * 2 + half);
      *reinterpret_cast<uint4*>(thi + off) = hi;
      *reinterpret_cast<uint4*>(tlo + off) = lo;
    }
  }
  for (int idx = tid; idx < 64 * 16; idx += kBwdThreads) {
    const int i = idx >> 4, m = idx & 15;
    woutc[idx] = m < NOUT ? p.out_w[m * 64 + i] : 0.f;
  }
  for (int idx = tid; idx < 64 * CZ; idx += kBwdThreads) {
    const int i = idx / CZ, q = idx % CZ;
    wzc[idx] = q < 3 * S ? p.w_ih[0][(int64_t)((q / S) * 64 + i) * ld0 + (q % S)] : 0.f;
  }
  if (tid == 0) {
    mbar_init(&bars->full[0], kEpiThreads);
    mbar_init(&bars->full[1], kEpiThreads);
    mbar_init(&bars->empty[0], 1);
    mbar_init(&bars->empty[1], 1);
    mbar_init(&bars->in0, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const int64_t ntiles = (p.B + kTileRows - 1) / kTileRows;

  // ---- MMA issue: chunk gc is issued by lane 0 of warp gc % 8 once all 256 threads have written it (the issue
  // order across warps is chained through the `full` barriers; the tensor pipe executes in issue order, so the
  // commit of a later chunk also covers the earlier ones)
  const uint32_t w0 = smem_u32(smem + L::OFF_W), a0 = smem_u32(smem + L::OFF_A);
  constexpr uint32_t ID64 = idesc_f16(64);
  // 9 MMAs: acc[128,64] (+)= A_chunk[slots] . W^T[K-groups of chunk c]
  auto issue = [&](uint32_t acc, uint32_t slot_base, int m, int c, bool n_is_nh, bool fresh) {
    const uint32_t a_hi = slot_base, a_lo = slot_base + kATileBytes;
    const uint32_t b_hi = w0 + (uint32_t)(m * 2) * kWTileBytes, b_lo = b_hi + kWTileBytes;
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
  uint32_t gc = 0;  // chunks produced so far (ring position / phases); uniform over the CTA

  {
    // ======================= gate-cotangent epilogue =============================================
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    uint8_t* a_ring = smem + L::OFF_A;
    uint32_t ph_in0 = 0, xb = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t b_raw = tile * kTileRows + row;
      const bool ok = b_raw < p.B;
      const float okf = ok ? 1.f : 0.f;
      const int64_t b = ok ? b_raw : p.B - 1;
      const float* st_tile = p.stash + tile * T * (int64_t)(NL * kStashSlots * 64 * kTileRows) + row;
      float* dg_tile = p.dg + tile * T * (int64_t)(NL * kDgSlots * 64 * kTileRows) + row;
      const float* gp_p = p.g_paths + b * (T + 1) * S;
      const float* gm_p = p.g_means + b * (int64_t)T * S;
      const float* ep_p = p.eps + b * (int64_t)T * S;
      const float* raw_p = p.raw + b * (int64_t)T * NTRIL;
      const float* gl_p = p.g_chol + b * (int64_t)T * S * S;

      // software pipeline: the stashed gates of the NEXT chunk (8 units x 5 values) and the per-step row inputs of
      // the NEXT step are loaded while the current ones are being processed
      // two register sets (even / odd chunks): a chunk's values are requested TWO chunks before they are used
      float pv[2][5][8];
      auto load_chunk = [&](float (&dst)[5][8], int tt, int kk, int cc) {
        const float* sk = st_tile + ((int64_t)tt * NL + kk) * (kStashSlots * 64 * kTileRows) + (cc * 16 + cg * 8) * kTileRows;
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
      // the chunk `ahead` positions after (tt, kk, cc) in processing order (c fastest, then layers down, then t-1)
      auto load_ahead = [&](float (&dst)[5][8], int tt, int kk, int cc, int ahead) {
        int lin = ((T - 1 - tt) * NL + (NL - 1 - kk)) * 4 + cc + ahead;
        const int t2 = T - 1 - lin / (4 * NL), k2 = NL - 1 - (lin / 4) % NL, c2 = lin % 4;
        if (t2 >= 0) load_chunk(dst, t2, k2, c2);
      };
      float n_gP[S], n_gM[S], n_ev[S], n_rd[S], n_gL[S * S];
      auto load_small = [&](int tt) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          n_gP[s] = gp_p[(tt + 1) * S + s];
          n_gM[s] = gm_p[tt * S + s];
          n_ev[s] = ep_p[tt * S + s];
          n_rd[s] = raw_p[tt * NTRIL + s * (s + 1) / 2 + s];
        }
#pragma unroll
        for (int q = 0; q < S * S; ++q) n_gL[q] = gl_p[tt * S * S + q];
      };
      load_small(T - 1);
      load_ahead(pv[0], T - 1, NL - 1, 0, 0);
      load_ahead(pv[1], T - 1, NL - 1, 0, 1);

      float dz[S];
#pragma unroll
      for (int s = 0; s < S; ++s) dz[s] = 0.f;
      float sc_prev[NL];  // 2^-(e + a_w) of the row scale used by step t+1 (per layer)
#pragma unroll
      for (int k = 0; k < NL; ++k) sc_prev[k] = 0.f;

      for (int t = T - 1; t >= 0; --t) {
        const bool first = t == T - 1;
        const uint32_t rpar = (uint32_t)(t & 1);
        // ---- cotangent of z_{t+1}: both threads of a row add the two partial sums in the same order
        if (!first) {
          named_bar_sync(1 + quad, 64);
#pragma unroll
          for (int s = 0; s < S; ++s) dz[s] += dzx[(0 * 128 + row) * S + s] + dzx[(1 * 128 + row) * S + s];
        }
        // ---- cotangent of the output projection (kernels/backward.py:300-334)
        float dout[NOUT];
        {
          float gP[S], gM[S], ev[S], rd[S], gL[S * S];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            gP[s] = n_gP[s] * okf;
            gM[s] = n_gM[s] * okf;
            ev[s] = n_ev[s];
            rd[s] = n_rd[s];
          }
#pragma unroll
          for (int q = 0; q < S * S; ++q) gL[q] = n_gL[q] * okf;
          if (t > 0) load_small(t - 1);
#pragma unroll
          for (int s = 0; s < S; ++s) dz[s] += gP[s];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            dout[s] = fmaf(dz[s], p.dt, gM[s]);
#pragma unroll
            for (int j = 0; j <= s; ++j) {
              float d = fmaf(dz[s] * ev[j], p.sqrt_dt, gL[s * S + j]);
              if (j == s) d = (rd[s] >= VISDE_DIAG_MIN || d < 0.f) ? d : 0.f;  // primitives/bounds.py:20
              dout[S + s * (s + 1) / 2 + j] = d;
            }
          }
          if (cg == 0) {  // tiled [tile][t][16][128]; pad rows hold zeros
            float* o = p.dout + (tile * T + t) * (int64_t)(16 * kTileRows) + row;
#pragma unroll
            for (int m = 0; m < NOUT; ++m) o[m * kTileRows] = dout[m];
          }
        }
        float dzp[S];
#pragma unroll
        for (int s = 0; s < S; ++s) dzp[s] = 0.f;
        float sc_in = 0.f;  // scale of dh_in0 (written by layer 1 of this step)

#pragma unroll
        for (int k = NL - 1; k >= 0; --k) {
          if (NL == 2 && k == 0) {
            mbar_wait(&bars->in0, ph_in0);
            ph_in0 ^= 1;
            tc_fence_after();
          }
          if (NL == 1 && !first) {
            // dhc_0 of the previous step comes from its last chunk's MMAs (with two layers the ring waits of
            // layer 1 already cover them)
            mbar_wait(&bars->empty[(gc - 1) & 1], ((gc - 1) >> 1) & 1);
            tc_fence_after();
          }
          // ---------- pass 1: dh of this thread's 32 units, row maximum ----------
          float dh[kUPT];
          float mx = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j0 = c * 16 + cg * 8;
            uint32_t va[8], vd[8], vi[8];
            if (!first) {
              tmem_ld8_nowait(tl + (uint32_t)(k * 2 + rpar) * 64 + j0, va);
              tmem_ld8_nowait(tl + DIR_COL + (uint32_t)k * 64 + j0, vd);
            }
            if (NL == 2 && k == 0) tmem_ld8_nowait(tl + IN0_COL + j0, vi);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float v = first ? 0.f : fmaf(sc_prev[k], __uint_as_float(va[q]), __uint_as_float(vd[q]));
              if (NL == 2 && k == 0) v = fmaf(sc_in, __uint_as_float(vi[q]), v);
              if (k == NL - 1) {
                const float* wc = woutc + (j0 + q) * 16;
#pragma unroll
                for (int m4 = 0; m4 < (NOUT + 3) / 4; ++m4) {
                  const float4 w4 = *reinterpret_cast<const float4*>(wc + 4 * m4);
                  v = fmaf(w4.x, dout[4 * m4], v);
                  if (4 * m4 + 1 < NOUT) v = fmaf(w4.y, dout[4 * m4 + 1 < NOUT ? 4 * m4 + 1 : 0], v);
                  if (4 * m4 + 2 < NOUT) v = fmaf(w4.z, dout[4 * m4 + 2 < NOUT ? 4 * m4 + 2 : 0], v);
                  if (4 * m4 + 3 < NOUT) v = fmaf(w4.w, dout[4 * m4 + 3 < NOUT ? 4 * m4 + 3 : 0], v);
                }
              }
              dh[c * 8 + q] = v;
              mx = fmaxf(mx, fabsf(v));
            }
          }
          // ---------- the two threads of the row agree on the power-of-two scale ----------
          maxb[(xb * 2 + cg) * 128 + row] = mx;
          named_bar_sync(1 + quad, 64);
          mx = fmaxf(mx, maxb[(xb * 2 + (cg ^ 1)) * 128 + row]);
          xb ^= 1;
          const int er = row_exp(mx);
          const float rs = exp2i(er);
          const float sc_this = exp2i(-(er + ew));

          // ---------- pass 2: gate cotangents, dg, direct term, A-operand chunks ----------
          float* dg_k = dg_tile + ((int64_t)t * NL + k) * (kDgSlots * 64 * kTileRows);
#pragma unroll
          for (int c = 0; c < 4; ++c, ++gc) {
            const int j0 = c * 16 + cg * 8;
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
              const int i = j0 + q;
              const float r = cr[q], u = cu[q], n = cn[q], nhh = cnh[q], hp = chp[q];
              const float dhv = dh[c * 8 + q];
              const float dnp = dhv * (1.f - u) * (1.f - n * n);
              const float dup = dhv * (hp - n) * u * (1.f - u);
              const float drp = dnp * nhh * r * (1.f - r);
              const float dnh = dnp * r;
              dirv[q] = __float_as_uint(dhv * u);
              dg_k[(0 * 64 + i) * kTileRows] = drp;
              dg_k[(1 * 64 + i) * kTileRows] = dup;
              dg_k[(2 * 64 + i) * kTileRows] = dnp;
              dg_k[(3 * 64 + i) * kTileRows] = dnh;
              if (k == 0) {
                const float* wz = wzc + i * CZ;
#pragma unroll
                for (int s = 0; s < S; ++s)
                  dzp[s] = fmaf(wz[s], drp, fmaf(wz[S + s], dup, fmaf(wz[2 * S + s], dnp, dzp[s])));
              }
              dr_[q] = drp * rs; du_[q] = dup * rs; dn_[q] = dnp * rs; dnh_[q] = dnh * rs;
            }
            tmem_st8(tl + DIR_COL + (uint32_t)k * 64 + j0, dirv);
            // ring slot: the MMAs that read its previous content (chunk gc - 2) must have completed
            const uint32_t slot = gc & 1;
            if (gc >= 2) mbar_wait(&bars->empty[slot], ((gc >> 1) - 1) & 1);
            uint8_t* ahi = a_ring + slot * SLOT_BYTES;
            uint4 hi, lo;
            split8(dr_, hi, lo);
            *reinterpret_cast<uint4*>(ahi + sw128(row, 0 + cg)) = hi;
            *reinterpret_cast<uint4*>(ahi + kATileBytes + sw128(row, 0 + cg)) = lo;
            split8(du_, hi, lo);
            *reinterpret_cast<uint4*>(ahi + sw128(row, 2 + cg)) = hi;
            *reinterpret_cast<uint4*>(ahi + kATileBytes + sw128(row, 2 + cg)) = lo;
            split8(dn_, hi, lo);
            *reinterpret_cast<uint4*>(ahi + sw128(row, 4 + cg)) = hi;
            *reinterpret_cast<uint4*>(ahi + kATileBytes + sw128(row, 4 + cg)) = lo;
            split8(dnh_, hi, lo);
            *reinterpret_cast<uint4*>(ahi + sw128(row, 6 + cg)) = hi;
            *reinterpret_cast<uint4*>(ahi + kATileBytes + sw128(row, 6 + cg)) = lo;
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->full[slot]);
            if (warp == (int)(gc & 7)) {
              mbar_wait(&bars->full[slot], (gc >> 1) & 1);
              tc_fence_after();
              if (lane == 0) {
                const uint32_t sb = a0 + slot * SLOT_BYTES;
                // accumulators written at step t are read at step t-1 (ping-pong by step parity)
                if (t > 0) issue(tmem + (uint32_t)(k * 2 + ((t & 1) ^ 1)) * 64, sb, k == 0 ? 0 : 2, c, true, c == 0);
                if (NL == 2 && k == 1) issue(tmem + IN0_COL, sb, 1, c, false, c == 0);
                umma_commit(&bars->empty[slot]);
                if (NL == 2 && k == 1 && c == 3) umma_commit(&bars->in0);
              }
              __syncwarp();
            }
          }
          tmem_st_wait();
          sc_prev[k] = sc_this;
          if (NL == 2 && k == 1) sc_in = sc_this;
        }
        // ---- this thread's share of d z_t through the state columns of W_ih_l0
#pragma unroll
        for (int s = 0; s < S; ++s) dzx[(cg * 128 + row) * S + s] = dzp[s];
      }
      // grad_x0 = d z_0 + g_paths[:, 0]
      named_bar_sync(1 + quad, 64);
      if (ok && cg == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s)
          p.grad_x0[b * S + s] = dz[s] + dzx[(0 * 128 + row) * S + s] + dzx[(1 * 128 + row) * S + s] + gp_p[s];
      }
      named_bar_sync(1 + quad, 64);  // dzx is rewritten by the next tile
    }
    // every MMA has completed before TMEM is released (in-order pipe: the last two commits cover all)
    if (gc >= 2) mbar_wait(&bars->empty[(gc - 2) & 1], ((gc - 2) >> 1) & 1);
    if (gc >= 1) mbar_wait(&bars->empty[(gc - 1) & 1], ((gc - 1) >> 1) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int NL, int S>
int launch_bwd_tc(const PathParams& p, cudaStream_t st) {
  const size_t smem = TcBwdSmem<NL, S>::bytes;
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_tc_kernel<NL, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (p.B + kTileRows - 1) / kTileRows;
  path_bwd_tc_kernel<NL, S><<<(unsigned)(ntiles < sms ? ntiles : sms), kBwdThreads, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

template <int NL>
int dispatch_s_tcb(const PathParams& p, cudaStream_t st) {
  switch (p.S) {
    case 1: return launch_bwd_tc<NL, 1>(p, st);
    case 2: return launch_bwd_tc<NL, 2>(p, st);
    case 3: return launch_bwd_tc<NL, 3>(p, st);
    case 4: return launch_bwd_tc<NL, 4>(p, st);
  }
  set_error("tensor-core recurrence: unsupported state dim %d", p.S);
  return VISDE_EINVAL;
}

}  // namespace

// p.stash, p.dg and p.dout are the row-fastest tiled buffers; p.grad_x0 per-trajectory
int launch_path_bwd_tc(const PathParams& p, cudaStream_t st) {
  if (p.NL == 1) return dispatch_s_tcb<1>(p, st);
  if (p.NL == 2) return dispatch_s_tcb<2>(p, st);
  set_error("tensor-core recurrence: unsupported num_layers %d", p.NL);
  return VISDE_EINVAL;
}

}  // namespace visde

// ------------------------------------------------------------------------------------------------
// Thin gradient pieces of the tensor-core family: everything that is a plain reduction of d_pre over
// (b, t) -- the bias gradients, the state columns of dW_ih_l0, dW_out / db_out and sum_t d_gi_l0 for the
// theta columns -- from the tiled dg / dout / stash in one coalesced pass (thread = trajectory row),
// per-tile partials, then a fixed-order sum over tiles (deterministic, no atomics).
// ------------------------------------------------------------------------------------------------
namespace visde {
namespace {

constexpr int kThinFPB = 32;  // dg features per block

// partial record of one tile: [F] bias sums, [192][S] dW_z, [NOUT][64] dW_out, [NOUT] db_out
__host__ __device__ inline int thin_part_floats(int NL, int S) {
  const int nout = S + S * (S + 1) / 2;
  return NL * kDgSlots * 64 + 192 * S + nout * 64 + nout;
}

template <int S>
__global__ void __launch_bounds__(256) tc_thin_kernel(const float* __restrict__ dg, const float* __restrict__ dout,
                                                      const float* __restrict__ stash, const float* __restrict__ paths,
                                                      int64_t B, int T, int NL, float* __restrict__ sdg,
                                                      float* __restrict__ part) {
  constexpr int NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL;
  const int F = NL * kDgSlots * 64;
  const int nfg = F / kThinFPB;
  const int64_t tb = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, rq = w & 3, par = w >> 2;
  const int row = rq * 32 + lane;
  const int64_t b_raw = tb * kTileRows + row;
  const bool ok = b_raw < B;
  const int64_t b = ok ? b_raw : B - 1;
  float* prec = part + tb * thin_part_floats(NL, S);
  __shared__ float red[2][4][16 * (1 + S)];

  if ((int)blockIdx.y < nfg) {
    // ---- bias sums, dW_z, sum_t d_gi_l0: 16 features per thread, lanes = rows
    const int f0 = blockIdx.y * kThinFPB + par * 16;
    const int64_t tstride = (int64_t)F * kTileRows;
    const float* src = dg + tb * T * tstride + (int64_t)f0 * kTileRows + row;
    const bool wz = f0 < 192;  // layer-0 slots r, u, n feed the state columns of W_ih_l0
    const float* zp = paths + b * (int64_t)(T + 1) * S;
    float acc[16], accz[16][S];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      acc[j] = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) accz[j][s] = 0.f;
    }
    for (int t = 0; t < T; ++t) {
      float z[S];
#pragma unroll
      for (int s = 0; s < S; ++s) z[s] = wz ? zp[t * S + s] : 0.f;
      const float* st = src + t * tstride;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = st[j * kTileRows];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        acc[j] += v[j];
#pragma unroll
        for (int s = 0; s < S; ++s) accz[j][s] = fmaf(v[j], z[s], accz[j][s]);
      }
    }
    if (wz && ok) {
#pragma unroll
      for (int j = 0; j < 16; ++j) sdg[b * 192 + f0 + j] = acc[j];
    }
    // sum over the 128 rows of the tile (pad rows hold exact zeros)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) red[par][rq][j] = a;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float c = accz[j][s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) red[par][rq][16 + j * S + s] = c;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * 16 * (1 + S); idx += blockDim.x) {
      const int pp = idx / (16 * (1 + S)), q = idx % (16 * (1 + S));
      const float a = (red[pp][0][q] + red[pp][1][q]) + (red[pp][2][q] + red[pp][3][q]);
      const int fb = blockIdx.y * kThinFPB + pp * 16;
      if (q < 16) prec[fb + q] = a;
      else if (fb < 192) prec[F + (fb + (q - 16) / S) * S + (q - 16) % S] = a;
    }
  } else {
    // ---- dW_out[m][i] = sum dout[m] h_top[i], db_out[m] = sum dout[m]: 8 units per block, 4 per thread
    const int ig = blockIdx.y - nfg;  // 0..7
    const int i0 = ig * 8 + par * 4;
    const int64_t sstride = (int64_t)NL * kStashSlots * 64 * kTileRows;
    const float* hsrc = stash + tb * T * sstride + ((int64_t)((NL - 1) * kStashSlots + kStashH) * 64 + i0) * kTileRows + row;
    const float* dsrc = dout + tb * T * (int64_t)(16 * kTileRows) + row;
    float acc[4][NOUT], accd[NOUT];
#pragma unroll
    for (int m = 0; m < NOUT; ++m) {
      accd[m] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][m] = 0.f;
    }
    for (int t = 0; t < T; ++t) {
      float dv[NOUT], h[4];
#pragma unroll
      for (int m = 0; m < NOUT; ++m) dv[m] = dsrc[((int64_t)t * 16 + m) * kTileRows];
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = hsrc[t * sstride + j * kTileRows];
#pragma unroll
      for (int m = 0; m < NOUT; ++m) {
        accd[m] += dv[m];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][m] = fmaf(dv[m], h[j], acc[j][m]);
      }
    }
    float* pw = prec + F + 192 * S;
    float* red2 = &red[0][0][0];  // [8 warps][4 * NOUT + NOUT]
    static_assert(8 * 5 * NOUT <= 2 * 4 * 16 * (1 + S), "reduction scratch too small");
#pragma unroll
    for (int m = 0; m < NOUT; ++m) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = acc[j][m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) red2[w * 5 * NOUT + j * NOUT + m] = a;
      }
      float d = accd[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (lane == 0) red2[w * 5 * NOUT + 4 * NOUT + m] = d;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * 5 * NOUT; idx += blockDim.x) {
      const int pp = idx / (5 * NOUT), q = idx % (5 * NOUT);
      const float* r0 = red2 + (pp * 4) * 5 * NOUT + q;
      const float a = (r0[0] + r0[5 * NOUT]) + (r0[2 * 5 * NOUT] + r0[3 * 5 * NOUT]);
      if (q < 4 * NOUT) pw[(q % NOUT) * 64 + ig * 8 + pp * 4 + q / NOUT] = a;
      else if (ig == 0 && pp == 0) pw[NOUT * 64 + (q - 4 * NOUT)] = a;
    }
  }
}

struct ThinReduceArgs {
  const float* part;
  int ntile, NL, S, n_out, ld0;
  float* b_ih[VISDE_MAX_LAYERS];
  float* b_hh[VISDE_MAX_LAYERS];
  float* w_ih0;
  float* out_w;
  float* out_b;
};
__global__ void tc_thin_reduce_kernel(ThinReduceArgs a) {
  const int total = thin_part_floats(a.NL, a.S);
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
    a.w_ih0[(int64_t)(off / a.S) * a.ld0 + off % a.S] = acc;  // row g*64+i, state column s
    return;
  }
  off -= 192 * a.S;
  if (off < a.n_out * 64) {
    a.out_w[off] = acc;
    return;
  }
  a.out_b[off - a.n_out * 64] = acc;
}

}  // namespace

size_t tc_thin_partial_floats(int64_t B, int NL, int S) {
  return (size_t)((B + kTileRows - 1) / kTileRows) * thin_part_floats(NL, S);
}

// dg / dout_tiled / stash: row-fastest tiled buffers of the tensor-core backward
int launch_tc_thin_grads(const PathParams& p, const float* dout_tiled, const visde_weight_grads* gw, float* partials,
                         size_t partial_floats, cudaStream_t st) {
  if (tc_thin_partial_floats(p.B, p.NL, p.S) > partial_floats) {
    set_error("tc thin gradients: workspace too small");
    return VISDE_EWORKSPACE;
  }
  const int64_t ntile = (p.B + kTileRows - 1) / kTileRows;
  const dim3 grid((unsigned)ntile, p.NL * kDgSlots * 64 / kThinFPB + 8);
  switch (p.S) {
    case 1: tc_thin_kernel<1><<<grid, 256, 0, st>>>(p.dg, dout_tiled, p.stash, p.paths, p.B, (int)p.T, p.NL, p.sdg, partials); break;
    case 2: tc_thin_kernel<2><<<grid, 256, 0, st>>>(p.dg, dout_tiled, p.stash, p.paths, p.B, (int)p.T, p.NL, p.sdg, partials); break;
    case 3: tc_thin_kernel<3><<<grid, 256, 0, st>>>(p.dg, dout_tiled, p.stash, p.paths, p.B, (int)p.T, p.NL, p.sdg, partials); break;
    case 4: tc_thin_kernel<4><<<grid, 256, 0, st>>>(p.dg, dout_tiled, p.stash, p.paths, p.B, (int)p.T, p.NL, p.sdg, partials); break;
    default: set_error("tc thin gradients: unsupported state dim %d", p.S); return VISDE_EINVAL;
  }
  VISDE_CUDA_CHECK(cudaGetLastError());
  ThinReduceArgs a{};
  a.part = partials;
  a.ntile = (int)ntile;
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
  const int total = thin_part_floats(p.NL, p.S);
  tc_thin_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace visde
