// Register-resident recurrence kernels for wide state spaces (4 < S <= 16, H <= 64, NL <= 2):
// K1s path_fwd_fasts, K2s path_bwd_fasts.  BASELINE config 5 (10-D stochastic Lorenz-96) lives here.
//
// Same skeleton as path_fast.cu -- one persistent CTA per trajectory slot, thread (i, ks) = hidden unit i,
// K-slice ks of 4, the three recurrent matrices in registers, FFMA2 dots, one barrier per layer -- but
// everything whose size grows with S (n_out = S + S(S+1)/2 reaches 152 rows at S = 16) moves from registers
// to shared memory, with S a run-time value:
//   * the state columns of W_ih_l0 and z_t: lane ks adds the terms s = ks, ks+4, ... to its partial
//     pre-activation, so the existing K-slice shuffle reduction sums them for free;
//   * the output projection: the 64 lane groups each take rows m = i, i+64, i+128 (weights from a bank-padded
//     shared copy, the h slice is already in registers), results meet in a shared vector, S threads run the
//     reparameterised Euler-Maruyama update (two more barriers per step than the small-S family);
//   * backward: d_out is produced by n_out threads from a shared ring of the per-step cotangents, W_out^T d_out
//     is split over the KS lanes, d z_t is 16 row ranges x S columns summed through shared memory.
// The bias gradients still accumulate in registers (per-CTA partials + fixed-order reduce); dW_ih_l0[:, :S],
// dW_out and db_out are left to the time-parallel GEMMs (api.cu).
#include "common.cuh"
#include "ptx.cuh"

namespace visde {
namespace {

constexpr int kHP = 64, kKS = 4, kSL = 16, kSLP = kSL + 4, kHB = kKS * kSLP;  // padded H-vector: 80 floats
constexpr int kThr = kHP * kKS;
constexpr int kWzPitch = 3 * kHP + 8;  // forward W_z copy: state column s at s * kWzPitch (4 consecutive s: disjoint banks)
constexpr int kWoPitch = 72;  // W_out^T rows in the backward: 4 consecutive rows hit 4 disjoint bank octets

__device__ __forceinline__ float ks_allreduce4(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ void load_slice16(const float* __restrict__ src, float2 (&dst)[8]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(src + 4 * q);
    dst[2 * q] = make_float2(v.x, v.y);
    dst[2 * q + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ float dot16(const float2 (&w)[8], const float2 (&x)[8]) {
  float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 8; q += 2) {
    fma2(a, w[q], x[q]);
    fma2(b, w[q + 1], x[q + 1]);
  }
  return (a.x + a.y) + (b.x + b.y);
}
__device__ __forceinline__ int padded16(int j) { return (j / kSL) * kSLP + (j % kSL); }

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int NL>
__global__ void __launch_bounds__(kThr, 1) path_fwd_fasts_kernel(PathParams p) {
  const int tid = threadIdx.x;
  const int i = tid / kKS, ks = tid % kKS;
  const int H = p.H, G = 3 * p.H, S = p.S, NOUT = p.n_out, NTRIL = p.n_tril, ld0 = p.S + p.C + p.P;
  const bool unit_ok = i < H;
  const float lead = ks == 0 ? 1.f : 0.f;
  const int npass = (NOUT + kHP - 1) / kHP;

  extern __shared__ __align__(16) float smem_fs[];
  float* hbuf = smem_fs;                        // [2][NL][kHB]
  float* wout_s = hbuf + 2 * NL * kHB;          // [npass * 64][kHB] rows of W_out as padded K-slices
  float* wz_s = wout_s + npass * kHP * kHB;     // [S][kWzPitch]: (g, unit) at g * kHP + unit
  float* bout_s = wz_s + S * kWzPitch;          // [NOUT]
  float* obuf = bout_s + NOUT;                  // [NOUT]
  float* zbuf = obuf + NOUT;                    // [S]
  float* epsbuf = zbuf + S;                     // [2][S]

  // ---- shared copies of the S-sized weights
  for (int idx = tid; idx < npass * kHP * kHB; idx += kThr) {
    const int m = idx / kHB, c = idx % kHB, k = (c / kSLP) * kSL + (c % kSLP);
    wout_s[idx] = (m < NOUT && (c % kSLP) < kSL && k < H) ? p.out_w[(int64_t)m * H + k] : 0.f;
  }
  for (int idx = tid; idx < S * 3 * kHP; idx += kThr) {
    const int u = idx % kHP, g = (idx / kHP) % 3, s = idx / (3 * kHP);
    wz_s[s * kWzPitch + g * kHP + u] = u < H ? p.w_ih[0][(int64_t)(g * H + u) * ld0 + s] : 0.f;
  }
  for (int m = tid; m < NOUT; m += kThr) bout_s[m] = p.out_b[m];

  // ---- recurrent weights into registers
  float2 whh[NL][3][kSL / 2];
  float2 wih[NL > 1 ? NL - 1 : 1][3][kSL / 2];
#pragma unroll
  for (int k = 0; k < NL; ++k)
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int q = 0; q < kSL; ++q) {
        const int kk = ks * kSL + q;
        const bool ok = unit_ok && kk < H;
        const float a = ok ? p.w_hh[k][(int64_t)(g * H + i) * H + kk] : 0.f;
        const float b = (ok && k > 0) ? p.w_ih[k][(int64_t)(g * H + i) * H + kk] : 0.f;
        if (q & 1) whh[k][g][q / 2].y = a; else whh[k][g][q / 2].x = a;
        if (k > 0) { if (q & 1) wih[k > 0 ? k - 1 : 0][g][q / 2].y = b; else wih[k > 0 ? k - 1 : 0][g][q / 2].x = b; }
      }
  float cb[NL][4];
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    float bir = 0.f, biu = 0.f, bin = 0.f;
    if (k > 0 && unit_ok) {
      bir = p.b_ih[k][i];
      biu = p.b_ih[k][H + i];
      bin = p.b_ih[k][2 * H + i];
    }
    cb[k][0] = unit_ok ? bir + p.b_hh[k][i] : 0.f;
    cb[k][1] = unit_ok ? biu + p.b_hh[k][H + i] : 0.f;
    cb[k][2] = bin;
    cb[k][3] = unit_ok ? p.b_hh[k][2 * H + i] : 0.f;
  }
  __syncthreads();

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    float gth[3] = {0.f, 0.f, 0.f};
    if (unit_ok) {
      for (int q = 0; q < p.P; ++q) {
        const float th = p.theta[b * p.P + q];
#pragma unroll
        for (int g = 0; g < 3; ++g) gth[g] += p.w_ih[0][(int64_t)(g * H + i) * ld0 + p.S + p.C + q] * th;
      }
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) gth[g] += cb[0][g];
    float hreg[NL], acc_hh[NL][3];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      hreg[k] = 0.f;
      acc_hh[k][0] = acc_hh[k][1] = acc_hh[k][2] = 0.f;
    }
    float eps_next = 0.f;
    if (tid < S) {
      const float v = p.x0[b * S + tid];
      zbuf[tid] = v;
      p.paths[b * (p.T + 1) * S + tid] = v;
      epsbuf[tid] = p.T > 0 ? p.eps[b * p.T * S + tid] : 0.f;  // parity 0 = step 0
    }
    const float* gi_p = p.gi_ctx + b * p.T * G + (unit_ok ? i : 0);
    float gi_cur[3] = {0.f, 0.f, 0.f};
    if (unit_ok && p.T > 0) {
#pragma unroll
      for (int g = 0; g < 3; ++g) gi_cur[g] = gi_p[g * H];
    }
    const bool st_on = p.stash != nullptr && unit_ok;
    float* st_lane = p.stash ? p.stash + b * p.T * (int64_t)(NL * kStashSlots * H) + ks * H + (unit_ok ? i : 0) : nullptr;
    // output pointers of this thread (one element per array per step), advanced by one row per step
    float* paths_o = p.paths + (b * (p.T + 1) + 1) * S + (tid < S ? tid : 0);
    float* means_o = p.means + b * p.T * S + (tid < S ? tid : 0);
    float* chol_o = p.chol + b * p.T * S * S + (tid < S * S ? tid : 0);
    float* raw_o = p.raw ? p.raw + b * p.T * NTRIL + (tid < NTRIL ? tid : 0) : nullptr;
    const int chol_r = tid / S, chol_c = tid % S;
    __syncthreads();  // zbuf / epsbuf of this trajectory visible

    for (int64_t t = 0; t < p.T; ++t) {
      const int par = (int)(t & 1);
      const bool has_next = t + 1 < p.T;
      float gi_nxt[3] = {0.f, 0.f, 0.f};
      gi_p += G;
      if (unit_ok && has_next) {
#pragma unroll
        for (int g = 0; g < 3; ++g) gi_nxt[g] = gi_p[g * H];
      }
      if (tid < S && has_next) eps_next = p.eps[(b * p.T + t + 1) * S + tid];

      float a_in[3] = {0.f, 0.f, 0.f};
      float2 hs[kSL / 2];
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        float pr, pu, pni, pnh;
        if (k == 0) {
          // lane ks adds the state columns s = ks, ks + 4, ...; the shuffle reduction below sums them
          float er = lead * (gi_cur[0] + gth[0]), eu = lead * (gi_cur[1] + gth[1]), en = lead * (gi_cur[2] + gth[2]);
#pragma unroll
          for (int q = 0; q < VISDE_MAX_STATE / kKS; ++q) {  // predicated and unrolled: the loads of all terms overlap
            const int s = ks + kKS * q;
            if (s < S) {
              const float zs = zbuf[s];
              const float* wz = wz_s + s * kWzPitch + i;
              er = fmaf(wz[0], zs, er);
              eu = fmaf(wz[kHP], zs, eu);
              en = fmaf(wz[2 * kHP], zs, en);
            }
          }
          pr = er + acc_hh[0][0];
          pu = eu + acc_hh[0][1];
          pni = en;
          pnh = fmaf(lead, cb[0][3], acc_hh[0][2]);
        } else {
          pr = fmaf(lead, cb[k][0], a_in[0] + acc_hh[k][0]);
          pu = fmaf(lead, cb[k][1], a_in[1] + acc_hh[k][1]);
          pni = fmaf(lead, cb[k][2], a_in[2]);
          pnh = fmaf(lead, cb[k][3], acc_hh[k][2]);
        }
        pr = ks_allreduce4(pr);
        pnh = ks_allreduce4(pnh);
        pni = ks_allreduce4(pni);
        pu = ks_allreduce4(pu);
        const float r = sigmoid_f(pr);
        const float n = tanh_f(fmaf(r, pnh, pni));
        const float u = sigmoid_f(pu);
        const float hn = unit_ok ? fmaf(u, hreg[k] - n, n) : 0.f;
        hreg[k] = hn;
        if (ks == 0) hbuf[(par * NL + k) * kHB + padded16(i)] = hn;
        {
          float v = r;
          v = ks == 1 ? u : v;
          v = ks == 2 ? n : v;
          v = ks == 3 ? pnh : v;
          if (st_on) st_lane[k * kStashSlots * H] = v;
          if (st_on && ks == 0) st_lane[k * kStashSlots * H + kStashH * H] = hn;
        }
        __syncthreads();
        load_slice16(&hbuf[(par * NL + k) * kHB + ks * kSLP], hs);
        if (k + 1 < NL) {
#pragma unroll
          for (int g = 0; g < 3; ++g) a_in[g] = dot16(wih[k + 1 < NL ? k : 0][g], hs);
        }
#pragma unroll
        for (int g = 0; g < 3; ++g) acc_hh[k][g] = dot16(whh[k][g], hs);
      }
      // ---- output projection: lane group i takes rows i, i + 64, ...
#pragma unroll
      for (int ps = 0; ps < 3; ++ps) {  // n_out <= 152 = 3 passes of 64 rows; unrolled so that the passes overlap
        if (ps < npass) {
          const int m = ps * kHP + i;
          float2 ws[kSL / 2];
          load_slice16(wout_s + m * kHB + ks * kSLP, ws);
          const float v = ks_allreduce4(dot16(ws, hs));
          if (ks == 0 && m < NOUT) obuf[m] = v + bout_s[m];
        }
      }
      __syncthreads();
      // ---- reparameterised Euler-Maruyama update: one state dimension per thread computes z_{t+1} (the only part on
      // the critical path); means / chol / raw leave after the barrier, one element per thread, straight from obuf
      // (first version: S threads each walked their row of L serially while seven warps waited at the barrier below --
      // 19 % of the kernel's stall samples.)  Now thread (s, j) = (tid / 16, tid % 16) forms L[s][j] eps[j] and the 16 lanes
      // of a row reduce with four shuffles: the whole CTA takes part and the chain is 2 LDS + 4 SHFL.
      {
        const int s = tid >> 4, j = tid & 15;
        float term = 0.f;
        if (s < S && j <= s) {
          const float raw = obuf[S + s * (s + 1) / 2 + j];
          term = (j == s ? fmaxf(raw, VISDE_DIAG_MIN) : raw) * epsbuf[par * S + j];
        }
        term += __shfl_xor_sync(0xffffffffu, term, 8);
        term += __shfl_xor_sync(0xffffffffu, term, 4);
        term += __shfl_xor_sync(0xffffffffu, term, 2);
        term += __shfl_xor_sync(0xffffffffu, term, 1);
        if (j == 0 && s < S) zbuf[s] = zbuf[s] + obuf[s] * p.dt + term * p.sqrt_dt;
        if (tid < S) epsbuf[(par ^ 1) * S + tid] = eps_next;
      }
      __syncthreads();
      {
        if (tid < S) {
          *paths_o = zbuf[tid];
          *means_o = obuf[tid];
        }
        if (tid < S * S) {  // S <= 16: one element of the S x S factor per thread
          float L = 0.f;
          if (chol_c <= chol_r) {
            const float raw = obuf[S + chol_r * (chol_r + 1) / 2 + chol_c];
            L = (chol_c == chol_r) ? fmaxf(raw, VISDE_DIAG_MIN) : raw;
          }
          *chol_o = L;
        }
        if (raw_o && tid < NTRIL) *raw_o = obuf[S + tid];
        paths_o += S;
        means_o += S;
        chol_o += S * S;
        if (raw_o) raw_o += NTRIL;
      }
      if (st_lane) st_lane += NL * kStashSlots * H;
#pragma unroll
      for (int g = 0; g < 3; ++g) gi_cur[g] = gi_nxt[g];
    }
    __syncthreads();  // obuf / zbuf reuse by the next trajectory
  }
}

// ---------------------------------------------------------------------------------------------
// backward (BPTT)
// ---------------------------------------------------------------------------------------------
// floats of one CTA's bias partial record
__host__ __device__ constexpr int fasts_part_floats(int NL, int H) { return NL * kDgSlots * H; }

template <int NL>
__global__ void __launch_bounds__(kThr, 1) path_bwd_fasts_kernel(PathParams p) {
  const int tid = threadIdx.x;
  const int i = tid / kKS, ks = tid % kKS;
  const int H = p.H, G = 3 * p.H, S = p.S, NOUT = p.n_out, NTRIL = p.n_tril, ld0 = p.S + p.C + p.P;
  const bool unit_ok = i < H;
  const int srow = (int)stash_row_floats(NL, H);
  // per-step row inputs staged in shared memory: gP[t+1], gM, eps, raw diag (S each), gL (S x S)
  const int O_GP = 0, O_GM = S, O_EPS = 2 * S, O_RAW = 3 * S, O_GL = 4 * S, SMALL = 4 * S + S * S;

  extern __shared__ __align__(16) float smem_bs[];
  float* dgb = smem_bs;                                   // [2][NL][4][kHB]
  float* sring = dgb + 2 * NL * kDgSlots * kHB;           // [4][NL * 5 * kHP]
  float* small = sring + 4 * NL * kStashSlots * kHP;      // [4][SMALL]: 3 rows live; 4 slots so that the slot index is a mask
  float* woT = small + 4 * ((SMALL + 3) / 4 * 4);         // [NOUT][kWoPitch]  W_out[m][i]
  float* wzr = woT + NOUT * kWoPitch;                     // [3 * kHP][16]     W_ih_l0[g*H+i][s] at [(g*64+i)*16 + s]
  float* doutb = wzr + 3 * kHP * 16;                      // [NOUT]
  float* dzb = doutb + ((NOUT + 3) / 4 * 4);              // [2][16]
  float* redz = dzb + 32;                                 // [8 warps][16]
  __shared__ __align__(8) uint64_t sbar[4];
  const int SMALLP = (SMALL + 3) / 4 * 4;

  for (int idx = tid; idx < NOUT * kWoPitch; idx += kThr) {
    const int m = idx / kWoPitch, u = idx % kWoPitch;
    woT[idx] = u < H ? p.out_w[(int64_t)m * H + u] : 0.f;
  }
  for (int idx = tid; idx < 3 * kHP * 16; idx += kThr) {
    const int s = idx % 16, u = (idx / 16) % kHP, g = idx / (16 * kHP);
    wzr[idx] = (s < S && u < H) ? p.w_ih[0][(int64_t)(g * H + u) * ld0 + s] : 0.f;
  }

  float2 whhT[NL][3][kSL / 2];
  float2 wihT[NL > 1 ? NL - 1 : 1][3][kSL / 2];
#pragma unroll
  for (int k = 0; k < NL; ++k)
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int q = 0; q < kSL; ++q) {
        const int kk = ks * kSL + q;
        const bool ok = unit_ok && kk < H;
        const float a = ok ? p.w_hh[k][(int64_t)(g * H + kk) * H + i] : 0.f;
        const float b = (ok && k > 0) ? p.w_ih[k][(int64_t)(g * H + kk) * H + i] : 0.f;
        if (q & 1) whhT[k][g][q / 2].y = a; else whhT[k][g][q / 2].x = a;
        if (k > 0) { if (q & 1) wihT[k > 0 ? k - 1 : 0][g][q / 2].y = b; else wihT[k > 0 ? k - 1 : 0][g][q / 2].x = b; }
      }
  float sb[NL];
#pragma unroll
  for (int k = 0; k < NL; ++k) sb[k] = 0.f;

  constexpr int NSR = 4;
  constexpr int kTmaThread = kThr - 32;
  const uint32_t row_bytes = (uint32_t)srow * 4u;
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < NSR; ++q) mbar_init(&sbar[q], 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t rows_issued = 0;
  const int sring_pitch = NL * kStashSlots * kHP;

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    float dhc[NL], sdg_lane = 0.f;
#pragma unroll
    for (int k = 0; k < NL; ++k) dhc[k] = 0.f;
    const int T = (int)p.T;
    const float* st_b = p.stash + b * p.T * (int64_t)srow;
    float* dg_p = p.dg + (b * p.T + (T - 1)) * (int64_t)(NL * kDgSlots * H) + ks * H + (unit_ok ? i : 0);

    // loader threads: element j (and j + kThr) of the small row of step r
    // loader threads: element j (and j + kThr) of the small row of step r sits at base[j] + r * stride[j]; the pointer
    // and the stride are formed once per trajectory, the time loop only steps the pointer back by one row
    auto small_ptr = [&](int j, int64_t* stride) -> const float* {
      if (j < O_GM) { *stride = S; return p.g_paths + (b * (p.T + 1) + 1) * S + j; }
      if (j < O_EPS) { *stride = S; return p.g_means + b * p.T * S + (j - O_GM); }
      if (j < O_RAW) { *stride = S; return p.eps + b * p.T * S + (j - O_EPS); }
      if (j < O_GL) { const int d = j - O_RAW; *stride = NTRIL; return p.raw + b * p.T * NTRIL + d * (d + 1) / 2 + d; }
      *stride = S * S;
      return p.g_chol + b * p.T * S * S + (j - O_GL);
    };
    float pend0 = 0.f, pend1 = 0.f;
    const float *src0 = nullptr, *src1 = nullptr;
    int64_t dec0 = 0, dec1 = 0;
    if (tid < SMALL) {
      src0 = small_ptr(tid, &dec0);
      if (T >= 1) small[((T - 1) & 3) * SMALLP + tid] = src0[(int64_t)(T - 1) * dec0];
      if (T >= 2) small[((T - 2) & 3) * SMALLP + tid] = src0[(int64_t)(T - 2) * dec0];
      if (T >= 3) pend0 = src0[(int64_t)(T - 3) * dec0];
      src0 += (int64_t)(T - 4) * dec0;  // next row to fetch (guarded by t >= 3 below)
    }
    if (tid + kThr < SMALL) {
      src1 = small_ptr(tid + kThr, &dec1);
      if (T >= 1) small[((T - 1) & 3) * SMALLP + tid + kThr] = src1[(int64_t)(T - 1) * dec1];
      if (T >= 2) small[((T - 2) & 3) * SMALLP + tid + kThr] = src1[(int64_t)(T - 2) * dec1];
      if (T >= 3) pend1 = src1[(int64_t)(T - 3) * dec1];
      src1 += (int64_t)(T - 4) * dec1;
    }
    if (tid < 32) dzb[tid] = 0.f;
    if (tid < 128) redz[tid] = 0.f;
    const uint32_t row0 = rows_issued;
    if (tid == kTmaThread) {
      for (int r = T - 1; r >= 0 && r >= T - 3; --r) {
        const uint32_t n = row0 + (uint32_t)(T - 1 - r);
        mbar_expect_tx(&sbar[n % NSR], row_bytes);
        bulk_load_1d(&sring[(n % NSR) * sring_pitch], st_b + (int64_t)r * srow, row_bytes, &sbar[n % NSR]);
      }
    }
    rows_issued += (uint32_t)T;
    __syncthreads();
    if (T > 0) mbar_wait(&sbar[row0 % NSR], (row0 / NSR) & 1);

    // row / column of the Cholesky entry behind output row m = tid (constant per thread: not recomputed every step)
    int out_r = tid, out_c = 0;
    if (tid >= S && tid < NOUT) {
      const int ti = tid - S;
      out_r = 0;
      while ((out_r + 1) * (out_r + 2) / 2 <= ti) ++out_r;
      out_c = ti - out_r * (out_r + 1) / 2;
    }
    for (int t = T - 1; t >= 0; --t) {
      const int par = t & 1;
      const uint32_t n_cur = row0 + (uint32_t)(T - 1 - t);
      const float* row_cur = &sring[(n_cur % NSR) * sring_pitch];
      const float* row_prev = &sring[((n_cur + 1) % NSR) * sring_pitch];
      if (tid == kTmaThread && t >= 3) {
        const uint32_t n = n_cur + 3;
        mbar_expect_tx(&sbar[n % NSR], row_bytes);
        bulk_load_1d(&sring[(n % NSR) * sring_pitch], st_b + (int64_t)(t - 3) * srow, row_bytes, &sbar[n % NSR]);
      }
      if (tid < SMALL) {
        if (t >= 2) small[((t - 2) & 3) * SMALLP + tid] = pend0;
        if (t >= 3) pend0 = *src0;
        src0 -= dec0;
      }
      if (tid + kThr < SMALL) {
        if (t >= 2) small[((t - 2) & 3) * SMALLP + tid + kThr] = pend1;
        if (t >= 3) pend1 = *src1;
        src1 -= dec1;
      }
      if (t >= 1) mbar_wait(&sbar[(n_cur + 1) % NSR], ((n_cur + 1) / NSR) & 1);
      const float* sm = small + (t & 3) * SMALLP;
      __syncthreads();  // the d z partial sums of step t+1 (redz) are complete

      // ---- cotangent of the output projection, one row per thread (kernels/backward.py:300-334); the thread
      // folds the pending d z partial sums of step t+1 into the d z it needs (double-buffered by step parity)
      if (tid < NOUT) {
        const int m = tid, r = out_r, c = out_c;
        float dz = dzb[par * 16 + r] + sm[O_GP + r];
#pragma unroll
        for (int w = 0; w < 8; ++w) dz += redz[w * 16 + r];
        float d;
        if (m < S) {
          d = fmaf(dz, p.dt, sm[O_GM + m]);
          dzb[(par ^ 1) * 16 + m] = dz;
        } else {
          d = fmaf(dz * sm[O_EPS + c], p.sqrt_dt, sm[O_GL + r * S + c]);
          if (r == c) d = (sm[O_RAW + r] >= VISDE_DIAG_MIN || d < 0.f) ? d : 0.f;  // primitives/bounds.py:20
        }
        doutb[m] = d;
        p.dout[(b * p.T + t) * NOUT + m] = d;
      }
      float c_r[NL], c_u[NL], c_n[NL], c_nhh[NL], c_hp[NL];
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        const int o = k * kStashSlots * H + (unit_ok ? i : 0);
        c_r[k] = row_cur[o + kStashR * H];
        c_u[k] = row_cur[o + kStashU * H];
        c_n[k] = row_cur[o + kStashN * H];
        c_nhh[k] = row_cur[o + kStashNhh * H];
        c_hp[k] = t > 0 ? row_prev[o + kStashH * H] : 0.f;
      }
      __syncthreads();  // d_out visible; redz / dzb of the previous step consumed

      // dh of the top layer: W_out^T d_out, rows m = ks, ks + 4, ... per lane
      float dh;
      {
        float a0 = 0.f, a1 = 0.f;  // two chains, unrolled: the loop is on the step's critical path
        const float* wc = woT + (unit_ok ? i : 0);
        int m = ks;
#pragma unroll 2
        for (; m + kKS < NOUT; m += 2 * kKS) {
          a0 = fmaf(wc[m * kWoPitch], doutb[m], a0);
          a1 = fmaf(wc[(m + kKS) * kWoPitch], doutb[m + kKS], a1);
        }
        if (m < NOUT) a0 = fmaf(wc[m * kWoPitch], doutb[m], a0);
        dh = dhc[NL - 1] + ks_allreduce4(a0 + a1);
      }
#pragma unroll
      for (int k = NL - 1; k >= 0; --k) {
        const float r = c_r[k], u = c_u[k], n = c_n[k];
        const float dnp = dh * (1.f - u) * (1.f - n * n);
        const float dup = dh * (c_hp[k] - n) * u * (1.f - u);
        const float drp = dnp * c_nhh[k] * r * (1.f - r);
        const float dnh = dnp * r;
        const float direct = dh * u;
        {
          float v = drp;
          v = ks == 1 ? dup : v;
          v = ks == 2 ? dnp : v;
          v = ks == 3 ? dnh : v;
          v = unit_ok ? v : 0.f;
          dgb[((par * NL + k) * kDgSlots + ks) * kHB + padded16(i)] = v;
          if (unit_ok) dg_p[k * kDgSlots * H] = v;
          sb[k] += v;
          if (k == 0) sdg_lane += v;
        }
        __syncthreads();
        const float* dk = dgb + (par * NL + k) * kDgSlots * kHB;
        float2 d0[kSL / 2], d1[kSL / 2], d2[kSL / 2], d3[kSL / 2];
        load_slice16(dk + 0 * kHB + ks * kSLP, d0);
        load_slice16(dk + 1 * kHB + ks * kSLP, d1);
        load_slice16(dk + 2 * kHB + ks * kSLP, d2);
        if (k > 0) {
          const float pb = ks_allreduce4(dot16(wihT[k > 0 ? k - 1 : 0][0], d0) + dot16(wihT[k > 0 ? k - 1 : 0][1], d1) +
                                         dot16(wihT[k > 0 ? k - 1 : 0][2], d2));
          dh = dhc[k > 0 ? k - 1 : 0] + pb;
        } else {
          // d z_t += W_ih_l0[:, :S]^T d_gi: thread = (column s = tid % 16, row range tid / 16 of 12 (gate, unit) rows)
          const int s = tid & 15, part = tid >> 4;
          float a = 0.f;
#pragma unroll
          for (int q = 0; q < 12; ++q) {
            const int gu = part * 12 + q, g = gu / kHP, uu = gu % kHP;
            a = fmaf(wzr[gu * 16 + s], dk[g * kHB + padded16(uu)], a);
          }
          a += __shfl_xor_sync(0xffffffffu, a, 16);
          if ((tid & 31) < 16) redz[(tid >> 5) * 16 + s] = a;
        }
        load_slice16(dk + 3 * kHB + ks * kSLP, d3);
        const float pc = ks_allreduce4(dot16(whhT[k][0], d0) + dot16(whhT[k][1], d1) + dot16(whhT[k][2], d3));
        dhc[k] = direct + pc;
      }
      dg_p -= NL * kDgSlots * H;
    }
    __syncthreads();  // last redz / dzb writes
    if (tid < S) {
      float dz = dzb[((0 & 1) ^ 1) * 16 + tid];  // written at step t = 0 (parity 0) into buffer 1
      if (T == 0) dz = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) dz += T > 0 ? redz[w * 16 + tid] : 0.f;
      p.grad_x0[b * S + tid] = dz + p.g_paths[b * (p.T + 1) * S + tid];
    }
    if (unit_ok && ks < 3) p.sdg[b * G + ks * H + i] = sdg_lane;
    __syncthreads();
  }
  if (p.cta_part && unit_ok) {
    float* part = p.cta_part + (int64_t)blockIdx.x * fasts_part_floats(NL, H);
#pragma unroll
    for (int k = 0; k < NL; ++k) part[(k * kDgSlots + ks) * H + i] = sb[k];
  }
}

struct FastsReduceArgs {
  const float* part;
  int ncta, NL, H;
  float* b_ih[VISDE_MAX_LAYERS];
  float* b_hh[VISDE_MAX_LAYERS];
};
__global__ void fasts_bias_reduce_kernel(FastsReduceArgs a) {
  const int total = fasts_part_floats(a.NL, a.H);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float acc = 0.f;
  for (int c = 0; c < a.ncta; ++c) acc += a.part[(int64_t)c * total + idx];
  const int H = a.H;
  const int k = idx / (kDgSlots * H), slot = (idx / H) % kDgSlots, i = idx % H;
  if (slot < 2) {
    a.b_ih[k][slot * H + i] = acc;
    a.b_hh[k][slot * H + i] = acc;
  } else if (slot == 2) {
    a.b_ih[k][2 * H + i] = acc;
  } else {
    a.b_hh[k][2 * H + i] = acc;
  }
}

int fasts_grid(int64_t B) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(B < sms ? B : sms);
}

size_t fasts_fwd_smem(const PathParams& p) {
  const int npass = (p.n_out + kHP - 1) / kHP;
  return sizeof(float) * ((size_t)2 * p.NL * kHB + (size_t)npass * kHP * kHB + (size_t)p.S * kWzPitch + 2 * (size_t)p.n_out + 3 * (size_t)p.S + 8);
}
size_t fasts_bwd_smem(const PathParams& p) {
  const int SMALL = 4 * p.S + p.S * p.S, SMALLP = (SMALL + 3) / 4 * 4;
  return sizeof(float) * ((size_t)2 * p.NL * kDgSlots * kHB + (size_t)4 * p.NL * kStashSlots * kHP + 4 * (size_t)SMALLP +
                          (size_t)p.n_out * kWoPitch + 3 * kHP * 16 + ((size_t)p.n_out + 3) / 4 * 4 + 32 + 128 + 8);
}

}  // namespace

bool fasts_supported(const PathParams& p) {
  return p.H <= 64 && p.H % 4 == 0 && p.NL <= 2 && p.S > 4 && p.S <= 16 && 4 * p.S + p.S * p.S <= 2 * kThr &&
         p.T * (int64_t)(p.NL * kStashSlots * p.H) < (int64_t(1) << 31);
}

int launch_path_fwd_fasts(const PathParams& p, cudaStream_t st) {
  const size_t smem = fasts_fwd_smem(p);
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_fwd_fasts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_fwd_fasts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_once.done(attr_dev);
  }
  if (p.NL == 1) path_fwd_fasts_kernel<1><<<fasts_grid(p.B), kThr, smem, st>>>(p);
  else path_fwd_fasts_kernel<2><<<fasts_grid(p.B), kThr, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

size_t fasts_partials_floats(int NL, int H) { return (size_t)256 * fasts_part_floats(NL, H); }

// p.cta_part must hold fasts_partials_floats(); writes the bias gradients
int launch_path_bwd_fasts(const PathParams& p, const visde_weight_grads* gw, cudaStream_t st) {
  const size_t smem = fasts_bwd_smem(p);
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_fasts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_fasts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_once.done(attr_dev);
  }
  const int grid = fasts_grid(p.B);
  if (p.NL == 1) path_bwd_fasts_kernel<1><<<grid, kThr, smem, st>>>(p);
  else path_bwd_fasts_kernel<2><<<grid, kThr, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  FastsReduceArgs a{};
  a.part = p.cta_part;
  a.ncta = grid;
  a.NL = p.NL;
  a.H = p.H;
  for (int k = 0; k < p.NL; ++k) {
    a.b_ih[k] = gw->b_ih[k];
    a.b_hh[k] = gw->b_hh[k];
  }
  const int total = fasts_part_floats(p.NL, p.H);
  fasts_bias_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace visde
