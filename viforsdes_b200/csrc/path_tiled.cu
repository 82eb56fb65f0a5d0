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

// trajectories per tile for a batch of B on this device (0 = use the one-trajectory-per-CTA family)
int tiled_batch_tile(int64_t B, bool force) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (B > (int64_t)4 * sms) return 8;
  if (B > sms || force) return 4;
  return 0;
}

int launch_path_fwd_tiled(const PathParams& p, int NB, cudaStream_t st) {
  if (p.NL == 1) return NB == 8 ? dispatch_s<1, 8>(p, st) : dispatch_s<1, 4>(p, st);
  if (p.NL == 2) return NB == 8 ? dispatch_s<2, 8>(p, st) : dispatch_s<2, 4>(p, st);
  set_error("tiled path: unsupported num_layers %d", p.NL);
  return VISDE_EINVAL;
}

}  // namespace visde
