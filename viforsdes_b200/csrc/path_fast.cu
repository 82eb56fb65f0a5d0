// Register-resident persistent recurrence kernels (the production family for H <= 64, NL <= 2,
// S <= 4): K1 path_fwd_fast, K2 path_bwd_fast.
//
// Design (B200-first, not a translation of kernels/forward.py / backward.py):
//   * one persistent CTA per trajectory slot, looping over all T steps; every recurrent weight
//     matrix (W_hh_l0, W_ih_l1, W_hh_l1 = 147 KB fp32 at H=64) lives in REGISTERS for the whole
//     launch: thread (i, ks) owns the K-slice ks of the three gate rows of hidden unit i
//     (forward) or of column i of the transposed matrices (backward);
//   * the hidden state of unit i stays in a register of its KS lanes; the only per-step exchange
//     is one H-float shared-memory vector per layer (parity double-buffered, one __syncthreads
//     per layer per step);
//   * the recurrent product W_hh_k h_k(t) needed by step t+1 is issued right after h_k(t) is
//     published, together with W_ih_{k+1} h_k(t): same operand, and it fills the pipeline while
//     the critical gate chain of the next layer waits on shuffles and MUFU;
//   * the context rows of W_ih_l0 (57 % of forward MACs) are hoisted into the time-parallel
//     GEMM K0 (gi_ctx); the kernel prefetches gi_ctx / eps one step ahead;
//   * the tiny output projection + Euler-Maruyama update is computed redundantly by every warp
//     (no extra barrier, z_{t+1} is in every thread's registers when layer 0 of step t+1 starts);
//   * backward emits d(pre-activations) [B,T,NL,4,H] and d(out) for the weight-gradient GEMMs
//     (K4) instead of accumulating weight gradients with atomics.
#include "common.cuh"

namespace visde {
namespace {

template <int KS>
__device__ __forceinline__ float ks_allreduce(float v) {
#pragma unroll
  for (int o = KS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_allreduce(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int SL>
__device__ __forceinline__ void load_slice(const float* __restrict__ src, float (&dst)[SL]) {
  static_assert(SL % 4 == 0, "slice must be float4-divisible");
#pragma unroll
  for (int q = 0; q < SL; q += 4) {
    float4 v = *reinterpret_cast<const float4*>(src + q);
    dst[q] = v.x;
    dst[q + 1] = v.y;
    dst[q + 2] = v.z;
    dst[q + 3] = v.w;
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int HP, int KS, int NL, int S>
__global__ void __launch_bounds__(HP* KS, 1) path_fwd_fast_kernel(PathParams p) {
  constexpr int SL = HP / KS, NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL, JC = HP / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = tid / KS, ks = tid % KS;
  const int H = p.H, G = 3 * p.H, ld0 = p.S + p.C + p.P;
  const bool unit_ok = i < H;
  const float lead = ks == 0 ? 1.f : 0.f;  // lane that adds the non-sliced terms before the reduce

  __shared__ __align__(16) float hbuf[2][NL][HP];

  // ---- weights into registers (once per CTA) ----
  float whh[NL][3][SL];
  float wih[NL > 1 ? NL - 1 : 1][3][SL];
#pragma unroll
  for (int k = 0; k < NL; ++k)
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int q = 0; q < SL; ++q) {
        const int kk = ks * SL + q;
        const bool ok = unit_ok && kk < H;
        whh[k][g][q] = ok ? p.w_hh[k][(int64_t)(g * H + i) * H + kk] : 0.f;
        if (k > 0) wih[k > 0 ? k - 1 : 0][g][q] = ok ? p.w_ih[k][(int64_t)(g * H + i) * H + kk] : 0.f;
      }
  float wz[3][S], cb[NL][4];  // cb: constant parts of (r, u, n_i, n_h) pre-activations
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int s = 0; s < S; ++s) wz[g][s] = unit_ok ? p.w_ih[0][(int64_t)(g * H + i) * ld0 + s] : 0.f;
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    float bir = 0.f, biu = 0.f, bin = 0.f;
    if (k > 0 && unit_ok) {
      bir = p.b_ih[k][i];
      biu = p.b_ih[k][H + i];
      bin = p.b_ih[k][2 * H + i];
    }  // layer 0: b_ih is folded into gi_ctx by K0
    cb[k][0] = unit_ok ? bir + p.b_hh[k][i] : 0.f;
    cb[k][1] = unit_ok ? biu + p.b_hh[k][H + i] : 0.f;
    cb[k][2] = bin;
    cb[k][3] = unit_ok ? p.b_hh[k][2 * H + i] : 0.f;
  }
  float wout[NOUT][JC], bout[NOUT];
#pragma unroll
  for (int m = 0; m < NOUT; ++m) {
    bout[m] = p.out_b[m];
#pragma unroll
    for (int c = 0; c < JC; ++c) {
      const int j = lane + 32 * c;
      wout[m][c] = j < H ? p.out_w[(int64_t)m * H + j] : 0.f;
    }
  }

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    // theta rows of W_ih_l0 applied once per trajectory (kernels/forward.py:157-175)
    float gth[3] = {0.f, 0.f, 0.f};
    if (unit_ok) {
      for (int q = 0; q < p.P; ++q) {
        const float th = p.theta[b * p.P + q];
#pragma unroll
        for (int g = 0; g < 3; ++g) gth[g] += p.w_ih[0][(int64_t)(g * H + i) * ld0 + p.S + p.C + q] * th;
      }
    }
    float z[S], hreg[NL], acc_hh[NL][3];
#pragma unroll
    for (int s = 0; s < S; ++s) z[s] = p.x0[b * S + s];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      hreg[k] = 0.f;
      acc_hh[k][0] = acc_hh[k][1] = acc_hh[k][2] = 0.f;
    }
    if (tid < S) p.paths[b * (p.T + 1) * S + tid] = p.x0[b * S + tid];

    const float* gi_base = p.gi_ctx + b * p.T * G;
    const float* eps_base = p.eps + b * p.T * S;
    float gi_cur[3] = {0.f, 0.f, 0.f}, eps_cur[S];
    if (unit_ok && p.T > 0) {
#pragma unroll
      for (int g = 0; g < 3; ++g) gi_cur[g] = gi_base[g * H + i];
    }
#pragma unroll
    for (int s = 0; s < S; ++s) eps_cur[s] = p.T > 0 ? eps_base[s] : 0.f;

    for (int64_t t = 0; t < p.T; ++t) {
      const int par = (int)(t & 1);
      const int64_t row = b * p.T + t;
      // prefetch next step's inputs
      float gi_nxt[3] = {0.f, 0.f, 0.f}, eps_nxt[S];
      const bool has_next = t + 1 < p.T;
      if (unit_ok && has_next) {
#pragma unroll
        for (int g = 0; g < 3; ++g) gi_nxt[g] = gi_base[(t + 1) * G + g * H + i];
      }
#pragma unroll
      for (int s = 0; s < S; ++s) eps_nxt[s] = has_next ? eps_base[(t + 1) * S + s] : 0.f;

      float a_in[3] = {0.f, 0.f, 0.f};  // W_ih_k h_{k-1}(t) slice partials for the layer being processed
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        float pr, pu, pni, pnh;
        if (k == 0) {
          float er = gi_cur[0] + gth[0] + cb[0][0], eu = gi_cur[1] + gth[1] + cb[0][1];
          float en = gi_cur[2] + gth[2] + cb[0][2];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            er = fmaf(wz[0][s], z[s], er);
            eu = fmaf(wz[1][s], z[s], eu);
            en = fmaf(wz[2][s], z[s], en);
          }
          pr = fmaf(lead, er, acc_hh[0][0]);
          pu = fmaf(lead, eu, acc_hh[0][1]);
          pni = lead * en;
          pnh = fmaf(lead, cb[0][3], acc_hh[0][2]);
        } else {
          pr = fmaf(lead, cb[k][0], a_in[0] + acc_hh[k][0]);
          pu = fmaf(lead, cb[k][1], a_in[1] + acc_hh[k][1]);
          pni = fmaf(lead, cb[k][2], a_in[2]);
          pnh = fmaf(lead, cb[k][3], acc_hh[k][2]);
        }
        pr = ks_allreduce<KS>(pr);
        pu = ks_allreduce<KS>(pu);
        pni = ks_allreduce<KS>(pni);
        pnh = ks_allreduce<KS>(pnh);
        const float r = sigmoid_f(pr), u = sigmoid_f(pu);
        const float n = tanh_f(fmaf(r, pnh, pni));
        const float hn = unit_ok ? fmaf(u, hreg[k] - n, n) : 0.f;  // (1-u) n + u h
        hreg[k] = hn;
        if (ks == 0) hbuf[par][k][i] = hn;
        if (p.stash && unit_ok) {
          float* st = p.stash + (row * NL + k) * (int64_t)(kStashSlots * H);
          // spread the five stores over the KS lanes of the unit (KS >= 4)
          if (ks == 0) { st[kStashR * H + i] = r; st[kStashH * H + i] = hn; }
          if (ks == 1) st[kStashU * H + i] = u;
          if (ks == 2) st[kStashN * H + i] = n;
          if (ks == 3) st[kStashNhh * H + i] = pnh;
        }
        __syncthreads();
        float hs[SL];
        load_slice<SL>(&hbuf[par][k][ks * SL], hs);
        if (k + 1 < NL) {
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            float a = 0.f;
#pragma unroll
            for (int q = 0; q < SL; ++q) a = fmaf(wih[k + 1 < NL ? k : 0][g][q], hs[q], a);
            a_in[g] = a;
          }
        }
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          float a = 0.f;
#pragma unroll
          for (int q = 0; q < SL; ++q) a = fmaf(whh[k][g][q], hs[q], a);
          acc_hh[k][g] = a;
        }
      }
      // output projection (every warp redundantly) + reparameterised Euler-Maruyama update
      float hv[JC], o[NOUT];
#pragma unroll
      for (int c = 0; c < JC; ++c) hv[c] = hbuf[par][NL - 1][lane + 32 * c];
#pragma unroll
      for (int m = 0; m < NOUT; ++m) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < JC; ++c) a = fmaf(wout[m][c], hv[c], a);
        o[m] = a;
      }
#pragma unroll
      for (int m = 0; m < NOUT; ++m) o[m] = warp_allreduce(o[m]) + bout[m];
      float zn[S], Lm[NTRIL];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j <= s; ++j) {
          const int ti = s * (s + 1) / 2 + j;
          const float raw = o[S + ti];
          const float L = (j == s) ? fmaxf(raw, VISDE_DIAG_MIN) : raw;
          Lm[ti] = L;
          acc = fmaf(L, eps_cur[j], acc);
        }
        zn[s] = z[s] + o[s] * p.dt + acc * p.sqrt_dt;
      }
      if (lane == 0) {
        if (warp == 0) {
#pragma unroll
          for (int s = 0; s < S; ++s) p.paths[(b * (p.T + 1) + t + 1) * S + s] = zn[s];
        } else if (warp == 1) {
#pragma unroll
          for (int s = 0; s < S; ++s) p.means[row * S + s] = o[s];
        } else if (warp == 2) {
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int j = 0; j < S; ++j) p.chol[(row * S + s) * S + j] = j <= s ? Lm[s * (s + 1) / 2 + j] : 0.f;
        } else if (warp == 3 && p.raw) {
#pragma unroll
          for (int ti = 0; ti < NTRIL; ++ti) p.raw[row * NTRIL + ti] = o[S + ti];
        }
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        z[s] = zn[s];
        eps_cur[s] = eps_nxt[s];
      }
#pragma unroll
      for (int g = 0; g < 3; ++g) gi_cur[g] = gi_nxt[g];
    }
    __syncthreads();  // hbuf reuse across trajectories
  }
}

// ---------------------------------------------------------------------------------------------
// backward (BPTT)
// ---------------------------------------------------------------------------------------------
template <int HP, int KS, int NL, int S>
__global__ void __launch_bounds__(HP* KS, 1) path_bwd_fast_kernel(PathParams p) {
  constexpr int SL = HP / KS, NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL, JC = HP / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = tid / KS, ks = tid % KS;
  const int H = p.H, G = 3 * p.H, ld0 = p.S + p.C + p.P;
  const bool unit_ok = i < H;
  const int64_t srow = stash_row_floats(NL, H);

  __shared__ __align__(16) float dgb[2][NL][kDgSlots][HP];

  // transposed weight slices: column i, rows ks*SL .. ks*SL+SL-1 of each gate block
  float whhT[NL][3][SL];
  float wihT[NL > 1 ? NL - 1 : 1][3][SL];
#pragma unroll
  for (int k = 0; k < NL; ++k)
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int q = 0; q < SL; ++q) {
        const int kk = ks * SL + q;
        const bool ok = unit_ok && kk < H;
        whhT[k][g][q] = ok ? p.w_hh[k][(int64_t)(g * H + kk) * H + i] : 0.f;
        if (k > 0) wihT[k > 0 ? k - 1 : 0][g][q] = ok ? p.w_ih[k][(int64_t)(g * H + kk) * H + i] : 0.f;
      }
  float woutc[NOUT];
#pragma unroll
  for (int m = 0; m < NOUT; ++m) woutc[m] = unit_ok ? p.out_w[(int64_t)m * H + i] : 0.f;
  // state columns of W_ih_l0 for the per-warp reduction of d z: lane owns rows j = lane + 32 c
  float wzl[3][S][JC];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int c = 0; c < JC; ++c) {
        const int j = lane + 32 * c;
        wzl[g][s][c] = j < H ? p.w_ih[0][(int64_t)(g * H + j) * ld0 + s] : 0.f;
      }

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    float dz[S], dhc[NL], sdg[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < S; ++s) dz[s] = 0.f;
#pragma unroll
    for (int k = 0; k < NL; ++k) dhc[k] = 0.f;

    // per-step inputs, software-prefetched one step ahead (reverse time)
    float c_gp[S], c_gm[S], c_gl[NTRIL], c_eps[S], c_rawd[S];
    float c_r[NL], c_u[NL], c_n[NL], c_nhh[NL], c_hp[NL];
    auto load_step = [&](int64_t t, float (&gp)[S], float (&gm)[S], float (&gl)[NTRIL], float (&ep)[S],
                         float (&rd)[S], float (&sr)[NL], float (&su)[NL], float (&sn)[NL],
                         float (&snh)[NL], float (&shp)[NL]) {
      const int64_t row = b * p.T + t;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        gp[s] = p.g_paths[(b * (p.T + 1) + t + 1) * S + s];
        gm[s] = p.g_means[row * S + s];
        ep[s] = p.eps[row * S + s];
        rd[s] = p.raw[row * NTRIL + s * (s + 1) / 2 + s];
#pragma unroll
        for (int j = 0; j <= s; ++j) gl[s * (s + 1) / 2 + j] = p.g_chol[(row * S + s) * S + j];
      }
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        sr[k] = su[k] = sn[k] = snh[k] = shp[k] = 0.f;
        if (unit_ok) {
          const float* st = p.stash + (row * NL + k) * (int64_t)(kStashSlots * H);
          sr[k] = st[kStashR * H + i];
          su[k] = st[kStashU * H + i];
          sn[k] = st[kStashN * H + i];
          snh[k] = st[kStashNhh * H + i];
          if (t > 0) shp[k] = (st - srow)[kStashH * H + i];
        }
      }
    };
    if (p.T > 0) load_step(p.T - 1, c_gp, c_gm, c_gl, c_eps, c_rawd, c_r, c_u, c_n, c_nhh, c_hp);

    for (int64_t t = p.T - 1; t >= 0; --t) {
      const int par = (int)(t & 1);
      const int64_t row = b * p.T + t;
      float n_gp[S], n_gm[S], n_gl[NTRIL], n_eps[S], n_rawd[S];
      float n_r[NL], n_u[NL], n_n[NL], n_nhh[NL], n_hp[NL];
      if (t > 0) load_step(t - 1, n_gp, n_gm, n_gl, n_eps, n_rawd, n_r, n_u, n_n, n_nhh, n_hp);

      // cotangent of the output projection
      float dout[NOUT];
#pragma unroll
      for (int s = 0; s < S; ++s) dz[s] += c_gp[s];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        dout[s] = fmaf(dz[s], p.dt, c_gm[s]);
#pragma unroll
        for (int j = 0; j <= s; ++j) {
          const int ti = s * (s + 1) / 2 + j;
          float d = fmaf(dz[s] * c_eps[j], p.sqrt_dt, c_gl[ti]);
          if (j == s && !(c_rawd[s] >= VISDE_DIAG_MIN || d < 0.f)) d = 0.f;  // primitives/bounds.py:20
          dout[S + ti] = d;
        }
      }
      if (warp == 0 && lane == 0) {
#pragma unroll
        for (int m = 0; m < NOUT; ++m) p.dout[row * NOUT + m] = dout[m];
      }
      float dh = dhc[NL - 1];
#pragma unroll
      for (int m = 0; m < NOUT; ++m) dh = fmaf(woutc[m], dout[m], dh);

#pragma unroll
      for (int k = NL - 1; k >= 0; --k) {
        const float r = c_r[k], u = c_u[k], n = c_n[k];
        const float dnp = dh * (1.f - u) * (1.f - n * n);
        const float dup = dh * (c_hp[k] - n) * u * (1.f - u);
        const float drp = dnp * c_nhh[k] * r * (1.f - r);
        const float dnh = dnp * r;
        const float direct = dh * u;
        if (unit_ok) {
          const float v = ks == 0 ? drp : ks == 1 ? dup : ks == 2 ? dnp : dnh;
          if (ks < 4) {
            dgb[par][k][ks][i] = v;
            p.dg[(row * NL + k) * (int64_t)(kDgSlots * H) + ks * H + i] = v;
          }
        } else if (ks < 4) {
          dgb[par][k][ks][i] = 0.f;
        }
        if (k == 0) {
          sdg[0] += drp;
          sdg[1] += dup;
          sdg[2] += dnp;
        }
        __syncthreads();
        float pc = 0.f, pb = 0.f;
        {
          float d0[SL], d1[SL];
          load_slice<SL>(&dgb[par][k][0][ks * SL], d0);
          load_slice<SL>(&dgb[par][k][1][ks * SL], d1);
#pragma unroll
          for (int q = 0; q < SL; ++q) {
            pc = fmaf(whhT[k][0][q], d0[q], pc);
            pc = fmaf(whhT[k][1][q], d1[q], pc);
            if (k > 0) {
              pb = fmaf(wihT[k > 0 ? k - 1 : 0][0][q], d0[q], pb);
              pb = fmaf(wihT[k > 0 ? k - 1 : 0][1][q], d1[q], pb);
            }
          }
        }
        {
          float d3[SL];
          load_slice<SL>(&dgb[par][k][3][ks * SL], d3);
#pragma unroll
          for (int q = 0; q < SL; ++q) pc = fmaf(whhT[k][2][q], d3[q], pc);
        }
        if (k > 0) {
          float d2[SL];
          load_slice<SL>(&dgb[par][k][2][ks * SL], d2);
#pragma unroll
          for (int q = 0; q < SL; ++q) pb = fmaf(wihT[k > 0 ? k - 1 : 0][2][q], d2[q], pb);
          pb = ks_allreduce<KS>(pb);
          dh = dhc[k > 0 ? k - 1 : 0] + pb;
        }
        pc = ks_allreduce<KS>(pc);
        dhc[k] = direct + pc;
        if (k == 0) {
          // d z_t += W_ih_l0[:, :S]^T d_gi (every warp redundantly; no barrier)
          float part[S];
#pragma unroll
          for (int s = 0; s < S; ++s) part[s] = 0.f;
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int c = 0; c < JC; ++c) {
              const float d = dgb[par][0][g][lane + 32 * c];
#pragma unroll
              for (int s = 0; s < S; ++s) part[s] = fmaf(wzl[g][s][c], d, part[s]);
            }
#pragma unroll
          for (int s = 0; s < S; ++s) dz[s] += warp_allreduce(part[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        c_gp[s] = n_gp[s];
        c_gm[s] = n_gm[s];
        c_eps[s] = n_eps[s];
        c_rawd[s] = n_rawd[s];
      }
#pragma unroll
      for (int ti = 0; ti < NTRIL; ++ti) c_gl[ti] = n_gl[ti];
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        c_r[k] = n_r[k];
        c_u[k] = n_u[k];
        c_n[k] = n_n[k];
        c_nhh[k] = n_nhh[k];
        c_hp[k] = n_hp[k];
      }
    }
    if (tid < S) {
      float v = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (s == tid) v = dz[s];
      p.grad_x0[b * S + tid] = v + p.g_paths[b * (p.T + 1) * S + tid];
    }
    if (unit_ok && ks < 3) {
      const float v = ks == 0 ? sdg[0] : ks == 1 ? sdg[1] : sdg[2];
      p.sdg[b * G + ks * H + i] = v;
    }
    __syncthreads();
  }
}

int fast_grid(int64_t B) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(B < sms ? B : sms);
}

template <int HP, int KS, int NL, int S>
int launch_fast(const PathParams& p, cudaStream_t st, bool bwd) {
  if (bwd)
    path_bwd_fast_kernel<HP, KS, NL, S><<<fast_grid(p.B), HP * KS, 0, st>>>(p);
  else
    path_fwd_fast_kernel<HP, KS, NL, S><<<fast_grid(p.B), HP * KS, 0, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

template <int HP, int KS, int NL>
int dispatch_s(const PathParams& p, cudaStream_t st, bool bwd) {
  switch (p.S) {
    case 1: return launch_fast<HP, KS, NL, 1>(p, st, bwd);
    case 2: return launch_fast<HP, KS, NL, 2>(p, st, bwd);
    case 3: return launch_fast<HP, KS, NL, 3>(p, st, bwd);
    case 4: return launch_fast<HP, KS, NL, 4>(p, st, bwd);
  }
  set_error("fast path: unsupported state dim %d", p.S);
  return VISDE_EINVAL;
}

template <int HP, int KS>
int dispatch_nl(const PathParams& p, cudaStream_t st, bool bwd) {
  if (p.NL == 1) return dispatch_s<HP, KS, 1>(p, st, bwd);
  if (p.NL == 2) return dispatch_s<HP, KS, 2>(p, st, bwd);
  set_error("fast path: unsupported num_layers %d", p.NL);
  return VISDE_EINVAL;
}

int dispatch_fast(const PathParams& p, cudaStream_t st, bool bwd) {
  if (p.H <= 32) return dispatch_nl<32, 8>(p, st, bwd);
  if (p.H <= 64) return dispatch_nl<64, 4>(p, st, bwd);
  set_error("fast path: unsupported hidden dim %d", p.H);
  return VISDE_EINVAL;
}

}  // namespace

bool fast_supported(const PathParams& p) { return p.H <= 64 && p.NL <= 2 && p.S <= 4; }
int launch_path_fwd_fast(const PathParams& p, cudaStream_t st) { return dispatch_fast(p, st, false); }
int launch_path_bwd_fast(const PathParams& p, cudaStream_t st) { return dispatch_fast(p, st, true); }

}  // namespace visde
