// Batch-tiled recurrence kernels for large batches (B > #SM): K1t path_fwd_tiled.
//
// The register-resident family (path_fast.cu) walks ONE trajectory per CTA and is bounded by the
// per-step dependency chain (~1 800 cycles per trajectory-step).  When there are more trajectories
// than SMs the same chain can carry NB = 4 or 8 trajectories at once: the recurrent weights move to
// shared memory (bank-padded rows, 184 KB at H = 64, NL = 2), every weight value fetched by a thread is
// reused for all NB trajectories of the tile (a register-blocked [3 gates x NB] micro-GEMM in packed
// FFMA2), the K-slice partials are combined by a reduce-scatter (each of the KS = 4 lanes of a unit
// ends up owning NB/4 trajectories, so gate math and stash stores are not replicated), and warp w
// runs the output projection + Euler-Maruyama update of trajectory w.  Per step the FP32 pipe does
// NB x the useful work for the same number of barriers, shuffles stages and MUFU round trips.
#include "common.cuh"

namespace visde {
namespace {

constexpr int kHP = 64, kKS = 4, kSL = 16, kSLP = 20;  // padded hidden, lanes per unit, K-slice, padded slice
constexpr int kRowP = kKS * kSLP;                       // padded row / vector length (80 floats)
constexpr int kTiledThreads = kHP * kKS;

__device__ __forceinline__ int padk(int k) { return (k / kSL) * kSLP + (k % kSL); }
// floats of one CTA's partial-sum record (same layout as path_fast.cu: biases, dW_ih_l0[:, :S], dW_out, db_out)
__host__ __device__ constexpr int fast_part_floats_tiled(int NL, int H, int S) {
  return NL * kDgSlots * H + 3 * S * H + (S + S * (S + 1) / 2) * H + (S + S * (S + 1) / 2);
}

// acc[g][n] (k-pair partial sums) += W[g rows of unit i, slice ks] . h[n][slice ks]
template <int NB>
__device__ __forceinline__ void tile_matvec(const float* __restrict__ wrow, const float* __restrict__ hsl,
                                            float2 (&acc)[3][NB]) {
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[g][n] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < kSL / 4; ++c) {
    float4 hv[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) hv[n] = *reinterpret_cast<const float4*>(hsl + n * kRowP + 4 * c);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 wv = *reinterpret_cast<const float4*>(wrow + g * kHP * kRowP + 4 * c);
      const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        fma2(acc[g][n], w01, make_float2(hv[n].x, hv[n].y));
        fma2(acc[g][n], w23, make_float2(hv[n].z, hv[n].w));
      }
    }
  }
}

// two matrices sharing the operand (W_ih_{k+1} and W_hh_k both multiply h_k): h chunks are loaded once
template <int NB>
__device__ __forceinline__ void tile_matvec2(const float* __restrict__ wrow_a, const float* __restrict__ wrow_b,
                                             const float* __restrict__ hsl, float2 (&acc_a)[3][NB],
                                             float2 (&acc_b)[3][NB]) {
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int n = 0; n < NB; ++n) acc_a[g][n] = acc_b[g][n] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < kSL / 4; ++c) {
    float4 hv[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) hv[n] = *reinterpret_cast<const float4*>(hsl + n * kRowP + 4 * c);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 wa = *reinterpret_cast<const float4*>(wrow_a + g * kHP * kRowP + 4 * c);
      const float4 wb = *reinterpret_cast<const float4*>(wrow_b + g * kHP * kRowP + 4 * c);
      const float2 a01 = make_float2(wa.x, wa.y), a23 = make_float2(wa.z, wa.w);
      const float2 b01 = make_float2(wb.x, wb.y), b23 = make_float2(wb.z, wb.w);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float2 h01 = make_float2(hv[n].x, hv[n].y), h23 = make_float2(hv[n].z, hv[n].w);
        fma2(acc_a[g][n], a01, h01);
        fma2(acc_a[g][n], a23, h23);
        fma2(acc_b[g][n], b01, h01);
        fma2(acc_b[g][n], b23, h23);
      }
    }
  }
}

// Reduce over the KS = 4 lanes of a unit and scatter: lane ks keeps the complete sums of trajectories
// ks * NB/4 .. ks * NB/4 + NB/4 - 1.  2 shuffle stages moving NB/2 + NB/4 values instead of 2 * NB.
template <int NB>
__device__ __forceinline__ void reduce_scatter(const float (&v)[NB], int ks, float (&out)[NB / 4]) {
  const bool hi2 = (ks & 2) != 0, hi1 = (ks & 1) != 0;
  float keep[NB / 2];
#pragma unroll
  for (int j = 0; j < NB / 2; ++j) {
    const float send = hi2 ? v[j] : v[NB / 2 + j];
    const float mine = hi2 ? v[NB / 2 + j] : v[j];
    keep[j] = mine + __shfl_xor_sync(0xffffffffu, send, 2);
  }
#pragma unroll
  for (int j = 0; j < NB / 4; ++j) {
    const float send = hi1 ? keep[j] : keep[NB / 4 + j];
    const float mine = hi1 ? keep[NB / 4 + j] : keep[j];
    out[j] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
  }
}

template <int NB>
__device__ __forceinline__ void reduce_scatter_gates(const float2 (&acc)[3][NB], int ks, float (&out)[3][NB / 4]) {
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    float v[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) v[n] = acc[g][n].x + acc[g][n].y;
    reduce_scatter<NB>(v, ks, out[g]);
  }
}

__device__ __forceinline__ float ks4_allreduce(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

template <int NL, int S, int NB>
__global__ void __launch_bounds__(kTiledThreads, 1) path_fwd_tiled_kernel(PathParams p) {
  constexpr int NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL, NPT = NB / 4, NMAT = 2 * NL - 1;
  constexpr int RPP = 32 / kKS, NP = (NOUT + RPP - 1) / RPP;
  extern __shared__ __align__(16) float smem_t[];
  float* Wsm = smem_t;                                  // [NMAT][3*HP][RowP]
  float* hb = Wsm + NMAT * 3 * kHP * kRowP;            // [NL][NB][RowP]
  float* zb = hb + NL * NB * kRowP;                     // [NB][4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = tid / kKS, ks = tid % kKS, grp = lane / kKS;
  const int H = p.H, G = 3 * p.H, ld0 = p.S + p.C + p.P;
  const bool unit_ok = i < H;
  const int nw = warp % NB;  // trajectory whose output projection / EM update this warp runs

  // ---- weights -> bank-padded shared memory (once per CTA); matrix order: W_hh_l0, W_ih_l1, W_hh_l1
  for (int m = 0; m < NMAT; ++m) {
    const float* src = m == 0 ? p.w_hh[0] : (m == 1 ? p.w_ih[1] : p.w_hh[1]);
    for (int idx = tid; idx < 3 * kHP * kHP; idx += kTiledThreads) {
      const int row = idx / kHP, k = idx % kHP, g = row / kHP, u = row % kHP;
      const float v = (u < H && k < H) ? src[(int64_t)(g * H + u) * H + k] : 0.f;
      Wsm[(m * 3 * kHP + row) * kRowP + padk(k)] = v;
    }
  }
  // pad columns are never read; zero the h / z vectors
  for (int idx = tid; idx < NL * NB * kRowP + NB * 4; idx += kTiledThreads) hb[idx] = 0.f;

  float wz[3][S], cb[NL][4];
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
    }
    cb[k][0] = unit_ok ? bir + p.b_hh[k][i] : 0.f;
    cb[k][1] = unit_ok ? biu + p.b_hh[k][H + i] : 0.f;
    cb[k][2] = bin;
    cb[k][3] = unit_ok ? p.b_hh[k][2 * H + i] : 0.f;
  }
  float2 wout[NP][kSL / 2];
  float bout[NOUT];
#pragma unroll
  for (int m = 0; m < NOUT; ++m) bout[m] = p.out_b[m];
#pragma unroll
  for (int ps = 0; ps < NP; ++ps)
#pragma unroll
    for (int q = 0; q < kSL; ++q) {
      const int m = ps * RPP + grp, kk = ks * kSL + q;
      const float a = (m < NOUT && kk < H) ? p.out_w[(int64_t)m * H + kk] : 0.f;
      if (q & 1) wout[ps][q / 2].y = a; else wout[ps][q / 2].x = a;
    }
  const float* wrow[NMAT];
#pragma unroll
  for (int m = 0; m < NMAT; ++m) wrow[m] = Wsm + (m * 3 * kHP + i) * kRowP + ks * kSLP;
  __syncthreads();

  const int64_t ngroups = (p.B + NB - 1) / NB;
  for (int64_t bg = blockIdx.x; bg < ngroups; bg += gridDim.x) {
    // trajectories owned by this lane (gate math, stash) and by this warp (out-proj, EM)
    int64_t bown[NPT];
    bool own_ok[NPT];
    float gth[NPT][3], hreg[NL][NPT], own_hh[NL][3][NPT];
    const float* gi_p[NPT];
    float gi_cur[NPT][3];
#pragma unroll
    for (int q = 0; q < NPT; ++q) {
      const int64_t b = bg * NB + ks * NPT + q;
      own_ok[q] = b < p.B;
      bown[q] = own_ok[q] ? b : p.B - 1;
#pragma unroll
      for (int g = 0; g < 3; ++g) gth[q][g] = cb[0][g];
      if (unit_ok) {
        for (int pp = 0; pp < p.P; ++pp) {
          const float th = p.theta[bown[q] * p.P + pp];
#pragma unroll
          for (int g = 0; g < 3; ++g) gth[q][g] += p.w_ih[0][(int64_t)(g * H + i) * ld0 + p.S + p.C + pp] * th;
        }
      }
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        hreg[k][q] = 0.f;
        own_hh[k][0][q] = own_hh[k][1][q] = own_hh[k][2][q] = 0.f;
      }
      gi_p[q] = p.gi_ctx + bown[q] * p.T * G + (unit_ok ? i : 0);
#pragma unroll
      for (int g = 0; g < 3; ++g) gi_cur[q][g] = (unit_ok && p.T > 0) ? gi_p[q][g * H] : 0.f;
    }
    const int64_t bw_raw = bg * NB + nw;
    const bool w_ok = bw_raw < p.B && warp < NB;  // warps beyond NB duplicate the math but never store
    const int64_t bw = bw_raw < p.B ? bw_raw : p.B - 1;
    float zw[S], eps_cur[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      zw[s] = p.x0[bw * S + s];
      eps_cur[s] = p.T > 0 ? p.eps[bw * p.T * S + s] : 0.f;
    }
    if (lane < S) {
      if (w_ok) p.paths[bw * (p.T + 1) * S + lane] = p.x0[bw * S + lane];
      float v = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) v = lane == s ? zw[s] : v;
      if (warp < NB) zb[nw * 4 + lane] = v;
    }
    const float* eps_p = p.eps + bw * p.T * S;
    float* paths_o = p.paths + (bw * (p.T + 1) + 1) * S;
    float* means_o = p.means + bw * p.T * S;
    float* chol_o = p.chol + bw * p.T * S * S;
    float* raw_o = p.raw ? p.raw + bw * p.T * NTRIL : nullptr;
    float* st_p[NPT];
#pragma unroll
    for (int q = 0; q < NPT; ++q)
      st_p[q] = p.stash ? p.stash + bown[q] * p.T * (int64_t)(NL * kStashSlots * H) + (unit_ok ? i : 0) : nullptr;
    __syncthreads();  // z of step 0 visible; previous tile's readers done

    for (int64_t t = 0; t < p.T; ++t) {
      const bool has_next = t + 1 < p.T;
      float gi_nxt[NPT][3], eps_nxt[S];
#pragma unroll
      for (int q = 0; q < NPT; ++q) {
        gi_p[q] += G;
#pragma unroll
        for (int g = 0; g < 3; ++g) gi_nxt[q][g] = (unit_ok && has_next) ? gi_p[q][g * H] : 0.f;
      }
      eps_p += S;
#pragma unroll
      for (int s = 0; s < S; ++s) eps_nxt[s] = has_next ? eps_p[s] : 0.f;

      float a_in[3][NPT];
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        // ---- gates of layer k for this lane's trajectories ----
#pragma unroll
        for (int q = 0; q < NPT; ++q) {
          float pr, pu, pni, pnh;
          if (k == 0) {
            const float* zq = zb + (ks * NPT + q) * 4;
            pr = gi_cur[q][0] + gth[q][0] + own_hh[0][0][q];
            pu = gi_cur[q][1] + gth[q][1] + own_hh[0][1][q];
            pni = gi_cur[q][2] + gth[q][2];
            pnh = own_hh[0][2][q] + cb[0][3];
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const float zs = zq[s];
              pr = fmaf(wz[0][s], zs, pr);
              pu = fmaf(wz[1][s], zs, pu);
              pni = fmaf(wz[2][s], zs, pni);
            }
          } else {
            pr = a_in[0][q] + own_hh[k][0][q] + cb[k][0];
            pu = a_in[1][q] + own_hh[k][1][q] + cb[k][1];
            pni = a_in[2][q] + cb[k][2];
            pnh = own_hh[k][2][q] + cb[k][3];
          }
          const float r = sigmoid_f(pr);
          const float n = tanh_f(fmaf(r, pnh, pni));
          const float u = sigmoid_f(pu);
          const float hn = unit_ok ? fmaf(u, hreg[k][q] - n, n) : 0.f;
          hreg[k][q] = hn;
          hb[(k * NB + ks * NPT + q) * kRowP + padk(i)] = hn;
          if (st_p[q] && unit_ok && own_ok[q]) {
            float* st = st_p[q] + k * kStashSlots * H;
            st[kStashR * H] = r;
            st[kStashU * H] = u;
            st[kStashN * H] = n;
            st[kStashNhh * H] = pnh;
            st[kStashH * H] = hn;
          }
        }
        __syncthreads();
        // ---- products with h_k(t) of all NB trajectories ----
        const float* hsl = hb + k * NB * kRowP + ks * kSLP;
        float2 acc[3][NB];
        if (k + 1 < NL) {
          // W_ih_{k+1} h_k (the next layer waits for it) and W_hh_k h_k (for step t+1) share the operand
          float2 acc2[3][NB];
          tile_matvec2<NB>(wrow[k + 1 < NL ? 2 * k + 1 : 0], wrow[k == 0 ? 0 : 2 * k], hsl, acc, acc2);
          reduce_scatter_gates<NB>(acc, ks, a_in);
          reduce_scatter_gates<NB>(acc2, ks, own_hh[k]);
        } else {
          tile_matvec<NB>(wrow[k == 0 ? 0 : 2 * k], hsl, acc);  // W_hh_k h_k for step t+1
          reduce_scatter_gates<NB>(acc, ks, own_hh[k]);
        }
      }
      // ---- output projection + reparameterised EM update of trajectory nw (one per warp) ----
      float o[NOUT];
      {
        float2 hs[kSL / 2];
        const float* hsrc = hb + ((NL - 1) * NB + nw) * kRowP + ks * kSLP;
#pragma unroll
        for (int q = 0; q < kSL / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(hsrc + 4 * q);
          hs[2 * q] = make_float2(v.x, v.y);
          hs[2 * q + 1] = make_float2(v.z, v.w);
        }
        float part[NP];
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
          float2 a = make_float2(0.f, 0.f), c = make_float2(0.f, 0.f);
#pragma unroll
          for (int q = 0; q < kSL / 2; q += 2) {
            fma2(a, wout[ps][q], hs[q]);
            fma2(c, wout[ps][q + 1], hs[q + 1]);
          }
          part[ps] = ks4_allreduce((a.x + a.y) + (c.x + c.y));
        }
#pragma unroll
        for (int m = 0; m < NOUT; ++m) o[m] = __shfl_sync(0xffffffffu, part[m / RPP], (m % RPP) * kKS) + bout[m];
      }
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
        zn[s] = zw[s] + o[s] * p.dt + acc * p.sqrt_dt;
      }
      if (w_ok) {
        if (lane == 0) {
#pragma unroll
          for (int s = 0; s < S; ++s) paths_o[s] = zn[s];
        } else if (lane == 8) {
#pragma unroll
          for (int s = 0; s < S; ++s) means_o[s] = o[s];
        } else if (lane == 16) {
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int j = 0; j < S; ++j) chol_o[s * S + j] = j <= s ? Lm[s * (s + 1) / 2 + j] : 0.f;
        } else if (lane == 24 && raw_o) {
#pragma unroll
          for (int ti = 0; ti < NTRIL; ++ti) raw_o[ti] = o[S + ti];
        }
      }
      if (lane < S && warp < NB) {
        float v = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) v = lane == s ? zn[s] : v;
        zb[nw * 4 + lane] = v;
      }
      paths_o += S;
      means_o += S;
      chol_o += S * S;
      if (raw_o) raw_o += NTRIL;
#pragma unroll
      for (int q = 0; q < NPT; ++q) {
        if (st_p[q]) st_p[q] += NL * kStashSlots * H;
#pragma unroll
        for (int g = 0; g < 3; ++g) gi_cur[q][g] = gi_nxt[q][g];
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        zw[s] = zn[s];
        eps_cur[s] = eps_nxt[s];
      }
      __syncthreads();  // z_{t+1} visible to every lane before the next step's layer-0 gates
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward (BPTT), same tiling: NB trajectories per CTA, TRANSPOSED recurrent weights in shared memory
// ---------------------------------------------------------------------------------------------
// Thread (j, ks) owns output unit j of the transposed products and the K-slice ks (16 gate rows of each gate block);
// after the reduce-scatter lane ks owns trajectories ks * NB/4 ... (gate derivatives, d_pre stores, thin-gradient
// accumulators); warp w additionally owns the scalar part of trajectory w % NB (d z, the output-projection cotangent,
// d z += W_z^T d_gi).  Emits the same d_pre / d_out / sum_t d_gi / per-CTA partial records as path_bwd_fast, so K3 / K4
// and fast_partials_reduce are shared.  Three barriers per step for NB trajectories (FAST: two per trajectory).

// acc_a[n] += sum_g W_ih^T[g] . d[n][slot g]   (slots r, u, n);  acc_b[n] += sum_g W_hh^T[g] . d[n][slot (r, u, n_hh)]
template <int NB>
__device__ __forceinline__ void bwd_matvec2(const float* __restrict__ wih, const float* __restrict__ whh,
                                            const float* __restrict__ dsl, float2 (&acc_a)[NB], float2 (&acc_b)[NB]) {
#pragma unroll
  for (int n = 0; n < NB; ++n) acc_a[n] = acc_b[n] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < kSL / 4; ++c) {
#pragma unroll
    for (int slot = 0; slot < kDgSlots; ++slot) {
      float4 dv[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) dv[n] = *reinterpret_cast<const float4*>(dsl + (n * kDgSlots + slot) * kRowP + 4 * c);
      if (slot < 3) {
        const float4 w = *reinterpret_cast<const float4*>(wih + slot * kHP * kRowP + 4 * c);
        const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
        for (int n = 0; n < NB; ++n) {
          fma2(acc_a[n], w01, make_float2(dv[n].x, dv[n].y));
          fma2(acc_a[n], w23, make_float2(dv[n].z, dv[n].w));
        }
      }
      if (slot != 2) {
        const float4 w = *reinterpret_cast<const float4*>(whh + (slot == 3 ? 2 : slot) * kHP * kRowP + 4 * c);
        const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
        for (int n = 0; n < NB; ++n) {
          fma2(acc_b[n], w01, make_float2(dv[n].x, dv[n].y));
          fma2(acc_b[n], w23, make_float2(dv[n].z, dv[n].w));
        }
      }
    }
  }
}

template <int NB>
__device__ __forceinline__ void bwd_matvec(const float* __restrict__ whh, const float* __restrict__ dsl, float2 (&acc_b)[NB]) {
#pragma unroll
  for (int n = 0; n < NB; ++n) acc_b[n] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < kSL / 4; ++c) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int slot = g == 2 ? 3 : g;
      const float4 w = *reinterpret_cast<const float4*>(whh + g * kHP * kRowP + 4 * c);
      const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float4 dv = *reinterpret_cast<const float4*>(dsl + (n * kDgSlots + slot) * kRowP + 4 * c);
        fma2(acc_b[n], w01, make_float2(dv.x, dv.y));
        fma2(acc_b[n], w23, make_float2(dv.z, dv.w));
      }
    }
  }
}

template <int NB>
__device__ __forceinline__ void reduce_scatter_sum(const float2 (&acc)[NB], int ks, float (&out)[NB / 4]) {
  float v[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) v[n] = acc[n].x + acc[n].y;
  reduce_scatter<NB>(v, ks, out);
}

constexpr int kSmallLd = 40;  // >= 5 S + S^2 for S <= 4, float4-aligned

template <int NL, int S, int NB>
__global__ void __launch_bounds__(kTiledThreads, 1) path_bwd_tiled_kernel(PathParams p) {
  constexpr int NTRIL = S * (S + 1) / 2, NOUT = S + NTRIL, NPT = NB / 4, NMAT = 2 * NL - 1;
  constexpr int RPP = 32 / kKS, NZ = 3 * S, NPZ = (NZ + RPP - 1) / RPP;
  constexpr int SMALL = 5 * S + S * S, SMALLP = (SMALL + 3) / 4 * 4;
  constexpr int O_GP = 0, O_GM = S, O_EPS = 2 * S, O_Z = 3 * S, O_RAW = 4 * S, O_GL = 5 * S;
  static_assert(SMALLP <= kSmallLd, "scalar row does not fit");
  extern __shared__ __align__(16) float smem_t[];
  float* WT = smem_t;                                   // [NMAT][3*HP][RowP] transposed: W_hh_l0, W_ih_l1, W_hh_l1
  float* dgb = WT + NMAT * 3 * kHP * kRowP;             // [NL][NB][4][RowP]
  float* doutb = dgb + NL * NB * kDgSlots * kRowP;      // [NB][16]
  float* smallb = doutb + NB * 16;                      // [warps][kSmallLd]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = tid / kKS, ks = tid % kKS, grp = lane / kKS;
  const int H = p.H, G = 3 * p.H, ld0 = p.S + p.C + p.P;
  const bool unit_ok = i < H;
  const int nw = warp % NB;
  const int T = (int)p.T;
  const int srow = (int)stash_row_floats(NL, H), dgrow = NL * kDgSlots * H;

  for (int m = 0; m < NMAT; ++m) {
    const float* src = m == 0 ? p.w_hh[0] : (m == 1 ? p.w_ih[1] : p.w_hh[1]);
    for (int idx = tid; idx < 3 * kHP * kHP; idx += kTiledThreads) {
      const int row = idx / kHP, j = idx % kHP, g = row / kHP, u = row % kHP;  // W[g*H + u][j] -> WT[g][j][u]
      const float v = (u < H && j < H) ? src[(int64_t)(g * H + u) * H + j] : 0.f;
      WT[(m * 3 * kHP + g * kHP + j) * kRowP + padk(u)] = v;
    }
  }
  for (int idx = tid; idx < NL * NB * kDgSlots * kRowP + NB * 16 + (kTiledThreads / 32) * kSmallLd; idx += kTiledThreads) dgb[idx] = 0.f;

  float woutc[NOUT];
#pragma unroll
  for (int m = 0; m < NOUT; ++m) woutc[m] = unit_ok ? p.out_w[(int64_t)m * H + i] : 0.f;
  float2 wzg[NPZ][kSL / 2];
#pragma unroll
  for (int ps = 0; ps < NPZ; ++ps)
#pragma unroll
    for (int q = 0; q < kSL; ++q) {
      const int item = ps * RPP + grp, kk = ks * kSL + q;
      const int gate = item / S, s = item % S;
      const float a = (item < NZ && kk < H) ? p.w_ih[0][(int64_t)(gate * H + kk) * ld0 + s] : 0.f;
      if (q & 1) wzg[ps][q / 2].y = a; else wzg[ps][q / 2].x = a;
    }
  const float* wrow[NMAT];
#pragma unroll
  for (int m = 0; m < NMAT; ++m) wrow[m] = WT + (m * 3 * kHP + i) * kRowP + ks * kSLP;

  // thin weight-gradient pieces over every step of every trajectory this lane owns
  float sb[NL][kDgSlots], wzacc[3][S], woacc[NOUT], dsum[NOUT];
#pragma unroll
  for (int k = 0; k < NL; ++k)
#pragma unroll
    for (int sl = 0; sl < kDgSlots; ++sl) sb[k][sl] = 0.f;
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int s = 0; s < S; ++s) wzacc[g][s] = 0.f;
#pragma unroll
  for (int m = 0; m < NOUT; ++m) woacc[m] = dsum[m] = 0.f;
  __syncthreads();

  const int64_t ngroups = (p.B + NB - 1) / NB;
  for (int64_t bg = blockIdx.x; bg < ngroups; bg += gridDim.x) {
    // ---- trajectories owned by this lane ----
    bool own_ok[NPT];
    const float* st_p[NPT];
    float* dg_p[NPT];
    float dhc[NL][NPT], sdg_acc[NPT][3], htop[NPT];
    float c_r[NL][NPT], c_u[NL][NPT], c_n[NL][NPT], c_nhh[NL][NPT], c_hp[NL][NPT];
#pragma unroll
    for (int q = 0; q < NPT; ++q) {
      const int64_t b = bg * NB + ks * NPT + q;
      own_ok[q] = b < p.B;
      const int64_t bo = own_ok[q] ? b : p.B - 1;
      st_p[q] = p.stash + bo * p.T * (int64_t)srow + (unit_ok ? i : 0);
      dg_p[q] = p.dg + bo * p.T * (int64_t)dgrow + (unit_ok ? i : 0);
      htop[q] = T > 0 ? st_p[q][(int64_t)(T - 1) * srow + ((NL - 1) * kStashSlots + kStashH) * H] : 0.f;
#pragma unroll
      for (int g = 0; g < 3; ++g) sdg_acc[q][g] = 0.f;
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        dhc[k][q] = 0.f;
        const float* row = st_p[q] + (int64_t)(T - 1) * srow + k * kStashSlots * H;
        c_r[k][q] = T > 0 ? row[kStashR * H] : 0.f;
        c_u[k][q] = T > 0 ? row[kStashU * H] : 0.f;
        c_n[k][q] = T > 0 ? row[kStashN * H] : 0.f;
        c_nhh[k][q] = T > 0 ? row[kStashNhh * H] : 0.f;
        c_hp[k][q] = T > 1 ? (row - srow)[kStashH * H] : 0.f;
      }
    }
    // ---- trajectory whose scalar part this warp owns ----
    const int64_t bw_raw = bg * NB + nw;
    const bool role_valid = bw_raw < p.B, w_ok = role_valid && warp < NB;
    const int64_t bw = role_valid ? bw_raw : p.B - 1;
    // lane j stages scalar j of the step (and scalar j + 32 when 5 S + S^2 > 32, i.e. S = 4), two steps ahead
    constexpr int NLD = SMALL > 32 ? 2 : 1;
    const float* src[NLD];
    int dec[NLD];
    bool ld_ok[NLD];  // rows of a trajectory beyond the batch read as zero: every cotangent vanishes
    float s_cur[NLD], s_nxt[NLD];
#pragma unroll
    for (int e = 0; e < NLD; ++e) {
      const int j = lane + 32 * e;
      src[e] = nullptr;
      dec[e] = 0;
      if (j < O_GM) { src[e] = p.g_paths + bw * (p.T + 1) * S + S + j; dec[e] = S; }
      else if (j < O_EPS) { src[e] = p.g_means + bw * p.T * S + (j - O_GM); dec[e] = S; }
      else if (j < O_Z) { src[e] = p.eps + bw * p.T * S + (j - O_EPS); dec[e] = S; }
      else if (j < O_RAW) { src[e] = p.paths + bw * (p.T + 1) * S + (j - O_Z); dec[e] = S; }
      else if (j < O_GL) { const int d = j - O_RAW; src[e] = p.raw + bw * p.T * NTRIL + d * (d + 1) / 2 + d; dec[e] = NTRIL; }
      else if (j < SMALL) { src[e] = p.g_chol + bw * p.T * S * S + (j - O_GL); dec[e] = S * S; }
      ld_ok[e] = j < SMALL && role_valid;
      s_cur[e] = (ld_ok[e] && T >= 1) ? src[e][(int64_t)(T - 1) * dec[e]] : 0.f;
      s_nxt[e] = (ld_ok[e] && T >= 2) ? src[e][(int64_t)(T - 2) * dec[e]] : 0.f;
    }
    float dz[S];
#pragma unroll
    for (int s = 0; s < S; ++s) dz[s] = 0.f;
    float* dout_p = p.dout + (bw * p.T + (T - 1)) * NOUT;
    float* sm_w = smallb + warp * kSmallLd;

    for (int t = T - 1; t >= 0; --t) {
      // ---- scalar part of trajectory nw: cotangent of the output projection ----
#pragma unroll
      for (int e = 0; e < NLD; ++e) {
        if (lane + 32 * e < SMALL) sm_w[lane + 32 * e] = s_cur[e];
        s_cur[e] = s_nxt[e];
        s_nxt[e] = (ld_ok[e] && t >= 2) ? src[e][(int64_t)(t - 2) * dec[e]] : 0.f;
      }
      __syncwarp();
      float sm[SMALLP];
#pragma unroll
      for (int q = 0; q < SMALLP / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(sm_w + 4 * q);
        sm[4 * q] = v.x; sm[4 * q + 1] = v.y; sm[4 * q + 2] = v.z; sm[4 * q + 3] = v.w;
      }
      float dout[NOUT];
#pragma unroll
      for (int s = 0; s < S; ++s) dz[s] += sm[O_GP + s];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        dout[s] = fmaf(dz[s], p.dt, sm[O_GM + s]);
#pragma unroll
        for (int j = 0; j <= s; ++j) {
          const int ti = s * (s + 1) / 2 + j;
          float d = fmaf(dz[s] * sm[O_EPS + j], p.sqrt_dt, sm[O_GL + s * S + j]);
          if (j == s) d = (sm[O_RAW + s] >= VISDE_DIAG_MIN || d < 0.f) ? d : 0.f;  // primitives/bounds.py:20
          dout[S + ti] = d;
        }
      }
      if (warp < NB && lane == 0) {
#pragma unroll
        for (int m = 0; m < NOUT; ++m) doutb[nw * 16 + m] = dout[m];
      }
      if (w_ok && lane == 8) {
#pragma unroll
        for (int m = 0; m < NOUT; ++m) dout_p[m] = dout[m];
      }
      dout_p -= NOUT;
      __syncthreads();  // B1: d_out of all NB trajectories (and their z_t) visible

      // stash of step t-1 for this lane's trajectories: in flight for the whole step
      float n_r[NL][NPT], n_u[NL][NPT], n_n[NL][NPT], n_nhh[NL][NPT], n_hp[NL][NPT];
#pragma unroll
      for (int q = 0; q < NPT; ++q)
#pragma unroll
        for (int k = 0; k < NL; ++k) {
          const float* row = st_p[q] + (int64_t)(t - 1) * srow + k * kStashSlots * H;
          n_r[k][q] = t >= 1 ? row[kStashR * H] : 0.f;
          n_u[k][q] = t >= 1 ? row[kStashU * H] : 0.f;
          n_n[k][q] = t >= 1 ? row[kStashN * H] : 0.f;
          n_nhh[k][q] = t >= 1 ? row[kStashNhh * H] : 0.f;
          n_hp[k][q] = t >= 2 ? (row - srow)[kStashH * H] : 0.f;
        }

      float dh[NPT], direct[NPT];
#pragma unroll
      for (int q = 0; q < NPT; ++q) {
        const float* dq = doutb + (ks * NPT + q) * 16;
        float a = dhc[NL - 1][q];
#pragma unroll
        for (int m = 0; m < NOUT; ++m) {
          const float d = dq[m];
          a = fmaf(woutc[m], d, a);
          woacc[m] = fmaf(d, htop[q], woacc[m]);
          dsum[m] += d;
        }
        dh[q] = a;
        htop[q] = c_hp[NL - 1][q];
      }

#pragma unroll
      for (int k = NL - 1; k >= 0; --k) {
#pragma unroll
        for (int q = 0; q < NPT; ++q) {
          const float r = c_r[k][q], u = c_u[k][q], n = c_n[k][q];
          float dnp = dh[q] * (1.f - u) * (1.f - n * n);
          float dup = dh[q] * (c_hp[k][q] - n) * u * (1.f - u);
          float drp = dnp * c_nhh[k][q] * r * (1.f - r);
          float dnh = dnp * r;
          direct[q] = dh[q] * u;
          if (!unit_ok) dnp = dup = drp = dnh = 0.f;
          float* db = dgb + ((k * NB + ks * NPT + q) * kDgSlots) * kRowP + padk(i);
          db[0] = drp;
          db[kRowP] = dup;
          db[2 * kRowP] = dnp;
          db[3 * kRowP] = dnh;
          if (unit_ok && own_ok[q]) {
            float* g = dg_p[q] + (int64_t)t * dgrow + k * kDgSlots * H;
            g[0] = drp;
            g[H] = dup;
            g[2 * H] = dnp;
            g[3 * H] = dnh;
          }
          sb[k][0] += drp;
          sb[k][1] += dup;
          sb[k][2] += dnp;
          sb[k][3] += dnh;
          if (k == 0) {
            sdg_acc[q][0] += drp;
            sdg_acc[q][1] += dup;
            sdg_acc[q][2] += dnp;
            const float* zq = smallb + (ks * NPT + q) * kSmallLd + O_Z;  // z_t of trajectory ks*NPT+q (warp index = trajectory)
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const float zs = zq[s];
              wzacc[0][s] = fmaf(drp, zs, wzacc[0][s]);
              wzacc[1][s] = fmaf(dup, zs, wzacc[1][s]);
              wzacc[2][s] = fmaf(dnp, zs, wzacc[2][s]);
            }
          }
        }
        __syncthreads();  // B2 / B3: d_pre of layer k of all NB trajectories visible
        const float* dsl = dgb + (k * NB * kDgSlots) * kRowP + ks * kSLP;
        if (k > 0) {
          float2 acc_a[NB], acc_b[NB];
          bwd_matvec2<NB>(wrow[2 * k - 1], wrow[2 * k], dsl, acc_a, acc_b);
          float pin[NPT], pc[NPT];
          reduce_scatter_sum<NB>(acc_a, ks, pin);
          reduce_scatter_sum<NB>(acc_b, ks, pc);
#pragma unroll
          for (int q = 0; q < NPT; ++q) {
            dh[q] = dhc[k > 0 ? k - 1 : 0][q] + pin[q];
            dhc[k][q] = direct[q] + pc[q];
          }
        } else {
          float2 acc_b[NB];
          bwd_matvec<NB>(wrow[0], dsl, acc_b);
          float pc[NPT];
          reduce_scatter_sum<NB>(acc_b, ks, pc);
#pragma unroll
          for (int q = 0; q < NPT; ++q) dhc[0][q] = direct[q] + pc[q];
          // d z_t += W_ih_l0[:, :S]^T d_gi of trajectory nw (KS-lane groups of this warp)
          float part[NPZ];
#pragma unroll
          for (int ps = 0; ps < NPZ; ++ps) {
            const int item = ps * RPP + grp;
            const int gate = item < NZ ? item / S : 0;
            const float* dsrc = dgb + (nw * kDgSlots + gate) * kRowP + ks * kSLP;  // layer 0
            float2 a = make_float2(0.f, 0.f), c = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < kSL / 4; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(dsrc + 4 * q);
              fma2(a, wzg[ps][2 * q], make_float2(v.x, v.y));
              fma2(c, wzg[ps][2 * q + 1], make_float2(v.z, v.w));
            }
            part[ps] = ks4_allreduce((a.x + a.y) + (c.x + c.y));
          }
#pragma unroll
          for (int s = 0; s < S; ++s)
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) {
              const int item = gate * S + s;
              dz[s] += __shfl_sync(0xffffffffu, part[item / RPP], (item % RPP) * kKS);
            }
        }
      }
#pragma unroll
      for (int q = 0; q < NPT; ++q)
#pragma unroll
        for (int k = 0; k < NL; ++k) {
          c_r[k][q] = n_r[k][q];
          c_u[k][q] = n_u[k][q];
          c_n[k][q] = n_n[k][q];
          c_nhh[k][q] = n_nhh[k][q];
          c_hp[k][q] = n_hp[k][q];
        }
    }
    if (w_ok && lane < S) {
      float v = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) v = lane == s ? dz[s] : v;
      p.grad_x0[bw * S + lane] = v + p.g_paths[bw * (p.T + 1) * S + lane];
    }
#pragma unroll
    for (int q = 0; q < NPT; ++q)
      if (unit_ok && own_ok[q]) {
        const int64_t b = bg * NB + ks * NPT + q;
#pragma unroll
        for (int g = 0; g < 3; ++g) p.sdg[b * G + g * H + i] = sdg_acc[q][g];
      }
    __syncthreads();  // shared buffers are reused by the next group of trajectories
  }

  // ---- per-CTA partial sums (layout of fast_part_floats, read by fast_partials_reduce_kernel) ----
  if (p.cta_part) {
    float* part = p.cta_part + (int64_t)blockIdx.x * fast_part_floats_tiled(NL, H, S);
#pragma unroll
    for (int k = 0; k < NL; ++k)
#pragma unroll
      for (int sl = 0; sl < kDgSlots; ++sl) {
        const float v = ks4_allreduce(sb[k][sl]);
        if (unit_ok && ks == 0) part[(k * kDgSlots + sl) * H + i] = v;
      }
    float* pz = part + NL * kDgSlots * H;
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float v = ks4_allreduce(wzacc[g][s]);
        if (unit_ok && ks == 0) pz[(g * S + s) * H + i] = v;
      }
    float* po = pz + 3 * S * H;
#pragma unroll
    for (int m = 0; m < NOUT; ++m) {
      const float v = ks4_allreduce(woacc[m]), dsm = ks4_allreduce(dsum[m]);
      if (unit_ok && ks == 0) po[m * H + i] = v;
      if (i == 0 && ks == 0) po[NOUT * H + m] = dsm;
    }
  }
}

size_t tiled_bwd_smem_bytes(int NL, int NB) {
  return sizeof(float) * ((size_t)(2 * NL - 1) * 3 * kHP * kRowP + (size_t)NL * NB * kDgSlots * kRowP + NB * 16 +
                          (kTiledThreads / 32) * kSmallLd);
}

size_t tiled_smem_bytes(int NL, int NB) {
  return sizeof(float) * ((size_t)(2 * NL - 1) * 3 * kHP * kRowP + (size_t)NL * NB * kRowP + NB * 4);
}

int tiled_grid(int64_t B, int NB) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t groups = (B + NB - 1) / NB;
  return (int)(groups < sms ? groups : sms);
}

template <int NL, int S, int NB>
int launch_tiled(const PathParams& p, cudaStream_t st) {
  const size_t smem = tiled_smem_bytes(NL, NB);
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_fwd_tiled_kernel<NL, S, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  path_fwd_tiled_kernel<NL, S, NB><<<tiled_grid(p.B, NB), kTiledThreads, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

template <int NL, int S, int NB>
int launch_tiled_bwd(const PathParams& p, cudaStream_t st) {
  const size_t smem = tiled_bwd_smem_bytes(NL, NB);
  static DeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.needed(&attr_dev)) {
    VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_tiled_kernel<NL, S, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_once.done(attr_dev);
  }
  path_bwd_tiled_kernel<NL, S, NB><<<tiled_grid(p.B, NB), kTiledThreads, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

template <int NL, int NB>
int dispatch_s_bwd(const PathParams& p, cudaStream_t st) {
  switch (p.S) {
    case 1: return launch_tiled_bwd<NL, 1, NB>(p, st);
    case 2: return launch_tiled_bwd<NL, 2, NB>(p, st);
    case 3: return launch_tiled_bwd<NL, 3, NB>(p, st);
    case 4: return launch_tiled_bwd<NL, 4, NB>(p, st);
  }
  set_error("tiled path: unsupported state dim %d", p.S);
  return VISDE_EINVAL;
}

template <int NL, int NB>
int dispatch_s(const PathParams& p, cudaStream_t st) {
  switch (p.S) {
    case 1: return launch_tiled<NL, 1, NB>(p, st);
    case 2: return launch_tiled<NL, 2, NB>(p, st);
    case 3: return launch_tiled<NL, 3, NB>(p, st);
    case 4: return launch_tiled<NL, 4, NB>(p, st);
  }
  set_error("tiled path: unsupported state dim %d", p.S);
  return VISDE_EINVAL;
}

}  // namespace

// trajectories per tile for a batch of B on this device (0 = use the one-trajectory-per-CTA family).
// Cost model in units of "one wave of one-trajectory CTAs", measured on B200 at H = 64, NL = 2, T = 100
// (tools/crossover_fast_tiled.sh, profiles/r1_ou_batch_sweep.md): a wave of 4-trajectory tiles costs 2.9 (forward) /
// 3.3 (backward) of them, a wave of 8-trajectory tiles 4.4 / 6.0; each family needs ceil(B / (tile * #SM)) waves.
int tiled_batch_tile(int64_t B, bool force, bool bwd) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  auto waves = [&](int tile) { return (double)((B + (int64_t)tile * sms - 1) / ((int64_t)tile * sms)); };
  const double c1 = waves(1), c4 = waves(4) * (bwd ? 3.32 : 2.9), c8 = waves(8) * (bwd ? 5.95 : 4.42);
  const int tiled = c8 < c4 ? 8 : 4;
  if (force) return tiled;
  return (c1 <= c4 && c1 <= c8) ? 0 : tiled;
}

int launch_path_fwd_tiled(const PathParams& p, int NB, cudaStream_t st) {
  if (p.NL == 1) return NB == 8 ? dispatch_s<1, 8>(p, st) : dispatch_s<1, 4>(p, st);
  if (p.NL == 2) return NB == 8 ? dispatch_s<2, 8>(p, st) : dispatch_s<2, 4>(p, st);
  set_error("tiled path: unsupported num_layers %d", p.NL);
  return VISDE_EINVAL;
}

// BPTT of the same tiling; *ncta receives the number of per-CTA partial records written to p.cta_part
int launch_path_bwd_tiled(const PathParams& p, int NB, int* ncta, cudaStream_t st) {
  *ncta = tiled_grid(p.B, NB);
  if (p.NL == 1) return NB == 8 ? dispatch_s_bwd<1, 8>(p, st) : dispatch_s_bwd<1, 4>(p, st);
  if (p.NL == 2) return NB == 8 ? dispatch_s_bwd<2, 8>(p, st) : dispatch_s_bwd<2, 4>(p, st);
  set_error("tiled path: unsupported num_layers %d", p.NL);
  return VISDE_EINVAL;
}

}  // namespace visde
