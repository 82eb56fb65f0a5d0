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
#include "path_tc.cuh"

namespace visde {
namespace {

constexpr int kBwdThreads = 256;
constexpr int kRowExp = 9;  // rows are scaled so that max|dh| 2^e is in [2^9, 2^10)

template <int S>
struct TcwBwdSmem {
  static constexpr int NOUT = S + S * (S + 1) / 2;
  static constexpr int CZ = 32;
  static constexpr int OFF_W1 = 0;                                // W_ih_l1^T hi, lo (resident)
  static constexpr int OFF_Y = kWImg;                             // time-shared: W_hh_l1^T / W_hh_l0^T
  static constexpr int OFF_A = 2 * kWImg;                         // ring [2][hi, lo][128][128 B]
  static constexpr int OFF_WOUT = OFF_A + 4 * kATileBytes;        // float [NOUT][64]: W_out[m][i]
  static constexpr int OFF_WZ = OFF_WOUT + NOUT * 64 * 4;         // float [64][CZ]: W_ih_l0[g*64+i][s] at [i][g*S+s]
  static constexpr int OFF_MAX = OFF_WZ + 64 * CZ * 4;            // float [2 buffers][2 cg][128]
  static constexpr int OFF_DZX = OFF_MAX + 2 * 2 * 128 * 4;       // float [2 cg][128][S]
  static constexpr int OFF_BAR = (OFF_DZX + 2 * 128 * S * 4 + 15) / 16 * 16;
  struct Bars {
    uint64_t full[2], empty[2], in0, wy, pro;
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

template <int S>
__global__ void __launch_bounds__(kBwdThreads, 1) path_bwd_tcw_kernel(PathParams p) {
  using L = TcwBwdSmem<S>;
  constexpr int NL = 2;
  constexpr int NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL, CZ = L::CZ, OF = tcw_out_feats(S), CF = tcw_cot_feats(S);
  static_assert(S > 4 && S <= kTcwMaxS && 3 * S <= CZ, "wide-state tensor-core recurrence: 4 < S <= 10");
  static_assert(L::bytes <= 227 * 1024, "shared memory budget");
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t IN0_COL = 256, DIR_COL = 320;
  constexpr int SLOT_BYTES = 2 * kATileBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw_tcwb[];
  uint8_t* smem = smem_raw_tcwb + ((1024u - (smem_u32(smem_raw_tcwb) & 1023u)) & 1023u);
  typename L::Bars* bars = reinterpret_cast<typename L::Bars*>(smem + L::OFF_BAR);
  float* woutm = reinterpret_cast<float*>(smem + L::OFF_WOUT);
  float* wzc = reinterpret_cast<float*>(smem + L::OFF_WZ);
  float* maxb = reinterpret_cast<float*>(smem + L::OFF_MAX);
  float* dzx = reinterpret_cast<float*>(smem + L::OFF_DZX);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld0 = S + p.C + p.P;
  const int T = (int)p.T;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.wimg);
  const int ew = reinterpret_cast<const int*>(img)[0];

  for (int idx = tid; idx < NOUT * 64; idx += kBwdThreads) woutm[idx] = p.out_w[idx];
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
    mbar_init(&bars->wy, 1);
    mbar_init(&bars->pro, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const int64_t ntiles = (p.B + kTileRows - 1) / kTileRows;
  if (tid == 0) {
    mbar_expect_tx(&bars->pro, kWImg);
    bulk_load_1d(smem + L::OFF_W1, img + kImgBwd0 + kWImg, kWImg, &bars->pro);
  }
  auto load_y = [&](int m) {  // thread 0: Y <- image of W_hh_l0^T (m = 0) / W_hh_l1^T (m = 2); the MMAs reading Y have completed
    mbar_expect_tx(&bars->wy, kWImg);
    bulk_load_1d(smem + L::OFF_Y, img + kImgBwd0 + (size_t)m * kWImg, kWImg, &bars->wy);
  };

  // ---- MMA issue: chunk gc is issued by lane 0 of warp gc % 8 once all 256 threads have written it
  const uint32_t w1 = smem_u32(smem + L::OFF_W1), wy = smem_u32(smem + L::OFF_Y), a0 = smem_u32(smem + L::OFF_A);
  constexpr uint32_t ID64 = idesc_f16(64);
  // 9 MMAs: acc[128,64] (+)= A_chunk[slots] . W^T[K-groups of chunk c]; wbase = hi tile of the transposed matrix
  auto issue = [&](uint32_t acc, uint32_t slot_base, uint32_t wbase, int c, bool n_is_nh, bool fresh) {
    const uint32_t a_hi = slot_base, a_lo = slot_base + kATileBytes;
    const uint32_t b_hi = wbase, b_lo = wbase + kWTileBytes;
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
    const int quad = warp & 3, cg = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(quad * 32) << 16);
    uint8_t* a_ring = smem + L::OFF_A;
    uint32_t ph_in0 = 0, xb = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t b_raw = tile * kTileRows + row;
      const bool ok = b_raw < p.B;
      const int64_t b = ok ? b_raw : p.B - 1;
      const float* st_tile = p.stash + tile * T * (int64_t)(NL * kStashSlots * 64 * kTileRows) + row;
      float* dg_tile = p.dg + tile * T * (int64_t)(NL * kDgSlots * 64 * kTileRows) + row;
      const float* ct_tile = p.ctile + tile * T * (int64_t)(CF * kTileRows) + row;  // pad rows hold zeros
      const float* ot_tile = p.otile + tile * T * (int64_t)(OF * kTileRows) + row;
      float* do_tile = p.dout + tile * T * (int64_t)(NOUT * kTileRows) + row;

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
      auto load_ahead = [&](float (&dst)[5][8], int tt, int kk, int cc, int ahead) {
        int lin = ((T - 1 - tt) * NL + (NL - 1 - kk)) * 4 + cc + ahead;
        const int t2 = T - 1 - lin / (4 * NL), k2 = NL - 1 - (lin / 4) % NL, c2 = lin % 4;
        if (t2 >= 0) load_chunk(dst, t2, k2, c2);
      };
      load_ahead(pv[0], T - 1, NL - 1, 0, 0);
      load_ahead(pv[1], T - 1, NL - 1, 0, 1);

      float dz[S];
#pragma unroll
      for (int s = 0; s < S; ++s) dz[s] = 0.f;
      float sc_prev[NL];
#pragma unroll
      for (int k = 0; k < NL; ++k) sc_prev[k] = 0.f;

      for (int t = T - 1; t >= 0; --t) {
        const bool first = t == T - 1;
        const uint32_t rpar = (uint32_t)(t & 1);
        const float* ct = ct_tile + (int64_t)t * (CF * kTileRows);
        // ---- Y <- W_hh_l1^T for this step's layer-1 phase (its carried products are issued for t >= 1 only): every MMA
        // issued so far has completed once the last chunk's commit has (in-order tensor pipe)
        if (tid == 0 && t >= 1) {
          if (gc >= 1) mbar_wait(&bars->empty[(gc - 1) & 1], ((gc - 1) >> 1) & 1);
          load_y(2);
        }
        // ---- cotangent of z_{t+1}: both threads of a row add the two partial sums in the same order
        if (!first) {
          named_bar_sync(1 + quad, 64);
#pragma unroll
          for (int s = 0; s < S; ++s) dz[s] += dzx[(0 * 128 + row) * S + s] + dzx[(1 * 128 + row) * S + s];
        }
        float ev[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          dz[s] += ct[s * kTileRows];                 // gP[t + 1]
          ev[s] = ct[(2 * S + S * S + s) * kTileRows];  // eps_t
        }
        float dzp[S];
#pragma unroll
        for (int s = 0; s < S; ++s) dzp[s] = 0.f;
        float sc_in = 0.f;

#pragma unroll
        for (int k = NL - 1; k >= 0; --k) {
          if (k == 0) {
            mbar_wait(&bars->in0, ph_in0);
            ph_in0 ^= 1;
            tc_fence_after();
            // every MMA of the layer-1 phase has completed (the in0 commit follows its last chunk): Y <- W_hh_l0^T
            if (tid == 0 && t >= 1) load_y(0);
          }
          // ---------- pass 1: dh of this thread's 32 units, row maximum ----------
          float dh[kUPT];
#pragma unroll
          for (int q = 0; q < kUPT; ++q) dh[q] = 0.f;
          if (k == NL - 1) {
            // cotangent of the output projection (kernels/backward.py:300-334), one entry at a time: d_out[m] is written
            // to the tiled buffer and contracted with W_out[m, this thread's 32 units]
            const float* otr = ot_tile + (int64_t)t * (OF * kTileRows);
            float* dor = do_tile + (int64_t)t * (NOUT * kTileRows);
            auto contract = [&](int m, float d) {
              if (cg == 0) dor[m * kTileRows] = d;
              const float* wr = woutm + m * 64 + cg * 8;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float4 wa = *reinterpret_cast<const float4*>(wr + c * 16);
                const float4 wb = *reinterpret_cast<const float4*>(wr + c * 16 + 4);
                dh[c * 8 + 0] = fmaf(wa.x, d, dh[c * 8 + 0]);
                dh[c * 8 + 1] = fmaf(wa.y, d, dh[c * 8 + 1]);
                dh[c * 8 + 2] = fmaf(wa.z, d, dh[c * 8 + 2]);
                dh[c * 8 + 3] = fmaf(wa.w, d, dh[c * 8 + 3]);
                dh[c * 8 + 4] = fmaf(wb.x, d, dh[c * 8 + 4]);
                dh[c * 8 + 5] = fmaf(wb.y, d, dh[c * 8 + 5]);
                dh[c * 8 + 6] = fmaf(wb.z, d, dh[c * 8 + 6]);
                dh[c * 8 + 7] = fmaf(wb.w, d, dh[c * 8 + 7]);
              }
            };
            float gl_cur[S], gl_nxt[S];  // lower-triangular row of gL, one row ahead
#pragma unroll
            for (int j = 0; j < 1; ++j) gl_cur[j] = ct[(2 * S + 0 * S + j) * kTileRows];
#pragma unroll
            for (int s = 0; s < S; ++s) {
              if (s + 1 < S) {
#pragma unroll
                for (int j = 0; j <= s + 1; ++j) gl_nxt[j] = ct[(2 * S + (s + 1) * S + j) * kTileRows];
              }
              const float gM = ct[(S + s) * kTileRows];
              const float rd = otr[(S + s * (s + 1) / 2 + s) * kTileRows];  // raw (unfloored) diagonal entry
              contract(s, fmaf(dz[s], p.dt, gM));
#pragma unroll
              for (int j = 0; j <= s; ++j) {
                float d = fmaf(dz[s] * ev[j], p.sqrt_dt, gl_cur[j]);
                if (j == s) d = (rd >= VISDE_DIAG_MIN || d < 0.f) ? d : 0.f;  // primitives/bounds.py:20
                contract(S + s * (s + 1) / 2 + j, d);
              }
#pragma unroll
              for (int j = 0; j < S; ++j) gl_cur[j] = gl_nxt[j];
            }
          }
          float mx = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j0 = c * 16 + cg * 8;
            uint32_t va[8], vd[8], vi[8];
            if (!first) {
              tmem_ld8_nowait(tl + (uint32_t)(k * 2 + rpar) * 64 + j0, va);
              tmem_ld8_nowait(tl + DIR_COL + (uint32_t)k * 64 + j0, vd);
            }
            if (k == 0) tmem_ld8_nowait(tl + IN0_COL + j0, vi);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float v = dh[c * 8 + q];
              if (!first) v += fmaf(sc_prev[k], __uint_as_float(va[q]), __uint_as_float(vd[q]));
              if (k == 0) v = fmaf(sc_in, __uint_as_float(vi[q]), v);
              dh[c * 8 + q] = v;
              mx = fmaxf(mx, fabsf(v));
            }
          }
          // ---------- the two threads of the row agree on the power-of-two scale ----------
          maxb[(xb * 2 + cg) * 128 + row] = mx;
          named_bar_sync(1 + quad, 64);
          mx = fmaxf(mx, maxb[(xb * 2 + (cg ^ 1)) * 128 + row]);
          xb ^= 1;
          const int er = row_exp_w(mx);
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
                // state columns of W_ih_l0: 3 S weights per unit, broadcast float4 reads
                const float* wz = wzc + i * CZ;
                float cc[CZ];
#pragma unroll
                for (int v = 0; v < (3 * S + 3) / 4; ++v) {
                  const float4 w4 = *reinterpret_cast<const float4*>(wz + 4 * v);
                  cc[4 * v] = w4.x; cc[4 * v + 1] = w4.y; cc[4 * v + 2] = w4.z; cc[4 * v + 3] = w4.w;
                }
#pragma unroll
                for (int s = 0; s < S; ++s)
                  dzp[s] = fmaf(cc[s], drp, fmaf(cc[S + s], dup, fmaf(cc[2 * S + s], dnp, dzp[s])));
              }
              dr_[q] = drp * rs; du_[q] = dup * rs; dn_[q] = dnp * rs; dnh_[q] = dnh * rs;
            }
            tmem_st8(tl + DIR_COL + (uint32_t)k * 64 + j0, dirv);
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
                if (t > 0) {
                  // Y holds W_hh_l1^T in the layer-1 phase (even completions of wy), W_hh_l0^T in the layer-0 phase (odd)
                  mbar_wait(&bars->wy, k == 1 ? 0u : 1u);
                  tc_fence_after();
                  issue(tmem + (uint32_t)(k * 2 + ((t & 1) ^ 1)) * 64, sb, wy, c, true, c == 0);
                }
                if (k == 1) {
                  if (gc < 8) {  // the first uses of the resident W_ih_l1^T tile: its prologue copy must have landed
                    mbar_wait(&bars->pro, 0);
                    tc_fence_after();
                  }
                  issue(tmem + IN0_COL, sb, w1, c, false, c == 0);
                }
                umma_commit(&bars->empty[slot]);
                if (k == 1 && c == 3) umma_commit(&bars->in0);
              }
              __syncwarp();
            }
          }
          tmem_st_wait();
          sc_prev[k] = sc_this;
          if (k == 1) sc_in = sc_this;
        }
#pragma unroll
        for (int s = 0; s < S; ++s) dzx[(cg * 128 + row) * S + s] = dzp[s];
      }
      // grad_x0 = d z_0 + g_paths[:, 0]
      named_bar_sync(1 + quad, 64);
      if (ok && cg == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s)
          p.grad_x0[b * S + s] = dz[s] + dzx[(0 * 128 + row) * S + s] + dzx[(1 * 128 + row) * S + s] + p.g_paths[b * (T + 1) * S + s];
      }
      named_bar_sync(1 + quad, 64);  // dzx is rewritten by the next tile
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
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_tcw_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (p.B + kTileRows - 1) / kTileRows;
  path_bwd_tcw_kernel<S><<<(unsigned)(ntiles < sms ? ntiles : sms), kBwdThreads, smem, st>>>(p);
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
constexpr int kTwUnits = 8;   // hidden units per thread in part B
constexpr int kTwM = 16;      // d_out entries per thread in part B

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
  for (int t = 0; t < T; ++t) {
    float z[S];
#pragma unroll
    for (int s = 0; s < S; ++s)
      z[s] = !wz ? 0.f : (t == 0 ? (ok ? paths[b * (int64_t)(T + 1) * S + s] : 0.f) : zrec[(int64_t)(t - 1) * (OF * kTileRows) + s * kTileRows]);
    const float* st = src + t * tstride;
    float v[kTwFeat];
#pragma unroll
    for (int j = 0; j < kTwFeat; ++j) v[j] = st[j * kTileRows];
#pragma unroll
    for (int j = 0; j < kTwFeat; ++j) {
      acc[j] += v[j];
#pragma unroll
      for (int s = 0; s < S; ++s) accz[j][s] = fmaf(v[j], z[s], accz[j][s]);
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

// grid (tile, unit group of 16 [2 par x 8 units], entry group of kTwM): acc[8 units][16 entries] per thread
template <int S>
__global__ void __launch_bounds__(256) tcw_thin_b_kernel(const float* __restrict__ dout, const float* __restrict__ stash, int T,
                                                         float* __restrict__ part) {
  constexpr int NL = 2, F = NL * kDgSlots * 64, NOUT = S + S * (S + 1) / 2;
  const int64_t tb = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, rq = w & 3, par = w >> 2;
  const int row = rq * 32 + lane;
  const int i0 = (blockIdx.y * 2 + par) * kTwUnits, m0 = blockIdx.z * kTwM;
  const int64_t sstride = (int64_t)NL * kStashSlots * 64 * kTileRows;
  const float* hsrc = stash + tb * T * sstride + ((int64_t)((NL - 1) * kStashSlots + kStashH) * 64 + i0) * kTileRows + row;
  const float* dsrc = dout + tb * T * (int64_t)(NOUT * kTileRows) + row;
  float acc[kTwUnits][kTwM], accd[kTwM];
#pragma unroll
  for (int m = 0; m < kTwM; ++m) {
    accd[m] = 0.f;
#pragma unroll
    for (int j = 0; j < kTwUnits; ++j) acc[j][m] = 0.f;
  }
  for (int t = 0; t < T; ++t) {
    float dv[kTwM], h[kTwUnits];
#pragma unroll
    for (int m = 0; m < kTwM; ++m) dv[m] = m0 + m < NOUT ? dsrc[((int64_t)t * NOUT + m0 + m) * kTileRows] : 0.f;
#pragma unroll
    for (int j = 0; j < kTwUnits; ++j) h[j] = hsrc[t * sstride + j * kTileRows];
#pragma unroll
    for (int m = 0; m < kTwM; ++m) {
      accd[m] += dv[m];
#pragma unroll
      for (int j = 0; j < kTwUnits; ++j) acc[j][m] = fmaf(dv[m], h[j], acc[j][m]);
    }
  }
  float* pw = part + tb * tcw_part_floats(NL, S) + F + 192 * S;
  __shared__ float red2[8][(kTwUnits + 1) * kTwM];
#pragma unroll
  for (int m = 0; m < kTwM; ++m) {
#pragma unroll
    for (int j = 0; j < kTwUnits; ++j) {
      float a = acc[j][m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) red2[w][j * kTwM + m] = a;
    }
    float d = accd[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) red2[w][kTwUnits * kTwM + m] = d;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * (kTwUnits + 1) * kTwM; idx += blockDim.x) {
    const int pp = idx / ((kTwUnits + 1) * kTwM), q = idx % ((kTwUnits + 1) * kTwM);
    const float a = (red2[pp * 4 + 0][q] + red2[pp * 4 + 1][q]) + (red2[pp * 4 + 2][q] + red2[pp * 4 + 3][q]);
    const int j = q / kTwM, m = m0 + q % kTwM;
    if (m >= NOUT) continue;
    if (j < kTwUnits) pw[m * 64 + (blockIdx.y * 2 + pp) * kTwUnits + j] = a;
    else if (blockIdx.y == 0 && pp == 0) pw[NOUT * 64 + m] = a;
  }
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
  off -= 192 * a.S;
  if (off < a.n_out * 64) {
    a.out_w[off] = acc;
    return;
  }
  a.out_b[off - a.n_out * 64] = acc;
}

template <int S>
int launch_thin_tcw(const PathParams& p, float* partials, cudaStream_t st) {
  const int64_t ntile = (p.B + kTileRows - 1) / kTileRows;
  constexpr int F = 2 * kDgSlots * 64, NOUT = S + S * (S + 1) / 2;
  tcw_thin_a_kernel<S><<<dim3((unsigned)ntile, F / (2 * kTwFeat)), 256, 0, st>>>(p.dg, p.otile, p.paths, p.B, (int)p.T, p.sdg, partials);
  VISDE_CUDA_CHECK(cudaGetLastError());
  tcw_thin_b_kernel<S><<<dim3((unsigned)ntile, 64 / (2 * kTwUnits), (NOUT + kTwM - 1) / kTwM), 256, 0, st>>>(p.dout, p.stash, (int)p.T,
                                                                                                         partials);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace

size_t tcw_image_bytes() { return (kImgBytes + 255) / 256 * 256; }

size_t tcw_thin_partial_floats(int64_t B, int NL, int S) {
  return (size_t)((B + kTileRows - 1) / kTileRows) * tcw_part_floats(NL, S);
}

// p.stash / p.dg / p.dout / p.otile / p.ctile are tiled buffers; p.wimg holds the backward images
int launch_path_bwd_tcw(const PathParams& p, cudaStream_t st) {
  // cotangent record [tile][t][3S + S*S][128]: gP[t+1] | gM | gL | eps
  const int S = p.S, CF = tcw_cot_feats(S);
  int rc = launch_tcw_tile(p.g_paths + S, p.B, p.T, S, (p.T + 1) * (int64_t)S, S, p.ctile, CF, 0, st);
  if (rc) return rc;
  if ((rc = launch_tcw_tile(p.g_means, p.B, p.T, S, p.T * (int64_t)S, S, p.ctile, CF, S, st))) return rc;
  if ((rc = launch_tcw_tile(p.g_chol, p.B, p.T, S * S, p.T * (int64_t)S * S, S * S, p.ctile, CF, 2 * S, st))) return rc;
  if ((rc = launch_tcw_tile(p.eps, p.B, p.T, S, p.T * (int64_t)S, S, p.ctile, CF, 2 * S + S * S, st))) return rc;
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
    case 5: rc = launch_thin_tcw<5>(p, partials, st); break;
    case 6: rc = launch_thin_tcw<6>(p, partials, st); break;
    case 7: rc = launch_thin_tcw<7>(p, partials, st); break;
    case 8: rc = launch_thin_tcw<8>(p, partials, st); break;
    case 9: rc = launch_thin_tcw<9>(p, partials, st); break;
    case 10: rc = launch_thin_tcw<10>(p, partials, st); break;
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
