// Time-parallel ELBO terms (forward and backward) -- K5/K6.
//
// Restates, as one pass over (b, tau) with tau in [0, T], the path-dependent terms of
// inference/evidence_lower_bound.py:28-56:
//   sde  = sum_t log N(x_{t+1}; x_t + f(x_t,th) dt, D(x_t,th) sqrt(dt))      (:42-44, :77-83)
//   gen  = sum_t log N(z_{t+1}; z_t + mu_t dt,     L_t sqrt(dt))             (:46-48)
//   jac  = sum_{t>=1} sum_{pos} logsigmoid(z_t)                              (:50, types.py:23-24)
//   obs  = sum_k log N(y_k; Hx[idx_k], var)                                  (:52-56, observations.py:52-74)
// with x = softplus(z) on positive dims (state_space.py:20-25; torch threshold 20), and f, D either
// the built-in Ornstein-Uhlenbeck / Lotka-Volterra functors (examples/*.py) or caller tensors.
// The Gaussian terms solve the triangular system in fp32 exactly like
// MultivariateNormal(scale_tril).log_prob, rather than using the analytic |eps|^2 shortcut, so the
// partial derivatives w.r.t. (z, mu, L) are the ones the reference's autograd produces.
// One CTA per trajectory; per-trajectory sums by block reduction (no atomics, deterministic).
#include "common.cuh"

namespace visde {
namespace {

constexpr int kElboThreads = 128;     // block size when there are many trajectories
constexpr int kElboMaxThreads = 512;  // few trajectories, long grids: more threads per trajectory (fewer serial round trips)
constexpr float kHalfLog2Pi = 0.91893853320467274178f;

__device__ __forceinline__ float softplus_f(float z) { return z > 20.f ? z : log1pf(expf(z)); }
__device__ __forceinline__ float softplus_grad_f(float z) { return z > 20.f ? 1.f : 1.f / (1.f + expf(-z)); }
__device__ __forceinline__ float logsigmoid_f(float z) { return fminf(z, 0.f) - log1pf(expf(-fabsf(z))); }

// y = A^{-1} r (forward substitution), lp = -1/2 |y|^2 - sum log A_ii - S/2 log 2pi,
// optionally w = A^{-T} y (back substitution).  A lower-triangular [SMAX][SMAX], S <= SMAX.
template <int SMAX>
__device__ __forceinline__ float gauss_solve(int S, const float (&A)[SMAX][SMAX], const float (&r)[SMAX],
                                             float (&y)[SMAX], float (&w)[SMAX], bool want_w) {
  float lp = 0.f;
#pragma unroll
  for (int i = 0; i < SMAX; ++i) {
    y[i] = 0.f;
    if (i < S) {
      float acc = r[i];
#pragma unroll
      for (int j = 0; j < i; ++j) acc -= A[i][j] * y[j];
      y[i] = acc / A[i][i];
      lp += -0.5f * y[i] * y[i] - logf(A[i][i]) - kHalfLog2Pi;
    }
  }
  if (want_w) {
#pragma unroll
    for (int i = SMAX - 1; i >= 0; --i) {
      w[i] = 0.f;
      if (i < S) {
        float acc = y[i];
#pragma unroll
        for (int j = i + 1; j < SMAX; ++j)
          if (j < S) acc -= A[j][i] * w[j];
        w[i] = acc / A[i][i];
      }
    }
  }
  return lp;
}

template <int SMAX>
__device__ __forceinline__ void load_vec(const float* p, int S, float (&v)[SMAX]) {
#pragma unroll
  for (int i = 0; i < SMAX; ++i) v[i] = i < S ? p[i] : 0.f;
}

// A thread's S x S factor block is S*S contiguous floats, the next thread's starts S*S floats later: scalar loads of the 55
// lower-triangular entries cost 55 instructions that each touch 32 scattered sectors per warp (elbo_bwd_kernel<10> sat at
// 1 TB/s on LSU issue).  When the block is a whole number of 16-byte words (S == SMAX, S*S % 4 == 0: S = 6, 8, 10, 12, 16)
// it is read with S*S/4 LDG.128 instead -- every sector fetched is consumed by two consecutive instructions of the same thread.
template <int SMAX>
__device__ __forceinline__ void load_tril_scaled(const float* p, int S, float scale, float (&A)[SMAX][SMAX], bool vec16 = false) {
  if (SMAX >= 6 && (SMAX * SMAX) % 4 == 0 && S == SMAX && vec16) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int q = 0; q < SMAX * SMAX / 4; ++q) {
      const float4 v = p4[q];
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = (4 * q + k) / SMAX, j = (4 * q + k) % SMAX;
        A[i][j] = j <= i ? e[k] * scale : 0.f;
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < SMAX; ++i)
#pragma unroll
    for (int j = 0; j < SMAX; ++j) A[i][j] = (i < S && j <= i) ? p[i * S + j] * scale : 0.f;
}

// cotangent of a factor block: G[i][j] = coef * (w[i] y[j] - [i == j] / A[i][i]) for j <= i, 0 above the diagonal; same
// 16-byte access rule as load_tril_scaled
template <int SMAX>
__device__ __forceinline__ void store_tril_grad(float* dst, int S, float coef, const float (&w)[SMAX], const float (&y)[SMAX],
                                                const float (&A)[SMAX][SMAX], bool vec16) {
  if (SMAX >= 6 && (SMAX * SMAX) % 4 == 0 && S == SMAX && vec16) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int q = 0; q < SMAX * SMAX / 4; ++q) {
      float e[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = (4 * q + k) / SMAX, j = (4 * q + k) % SMAX;
        e[k] = j <= i ? coef * (w[i] * y[j] - (i == j ? 1.f / A[i][i] : 0.f)) : 0.f;
      }
      d4[q] = make_float4(e[0], e[1], e[2], e[3]);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < SMAX; ++i)
#pragma unroll
    for (int j = 0; j < SMAX; ++j)
      if (i < S && j < S) dst[i * S + j] = j <= i ? coef * (w[i] * y[j] - (i == j ? 1.f / A[i][i] : 0.f)) : 0.f;
}

template <int SMAX>
__device__ __forceinline__ void to_state_vec(int S, uint32_t mask, const float (&z)[SMAX], float (&x)[SMAX]) {
#pragma unroll
  for (int i = 0; i < SMAX; ++i) x[i] = (i < S && ((mask >> i) & 1u)) ? softplus_f(z[i]) : z[i];
}

// Lotka-Volterra Cholesky factor of the diffusion matrix (examples/lotka_volterra.py:31-46)
struct LvDiff {
  float b11, b12, b22, L00, d00, L10, e, L11;
};
__device__ __forceinline__ LvDiff lv_diffusion(float u, float v, float t1, float t2, float t3) {
  LvDiff d;
  float uv = u * v;
  d.b11 = t1 * u + t2 * uv;
  d.b12 = -t2 * uv;
  d.b22 = t3 * v + t2 * uv;
  d.L00 = sqrtf(fmaxf(d.b11, 1e-6f));
  d.d00 = fmaxf(d.L00, 1e-6f);
  d.L10 = d.b12 / d.d00;
  d.e = d.b22 - d.L10 * d.L10;
  d.L11 = sqrtf(fmaxf(d.e, 1e-6f));
  return d;
}

// Builds the SDE transition's mean residual and scaled Cholesky factor for transition t of
// trajectory b:  r = x_next - (x_t + f dt),  A = D sqrt(dt).
template <int SMAX>
__device__ __forceinline__ void sde_transition(const ElboParams& p, int64_t b, int64_t t, const float* th,
                                               const float (&xt)[SMAX], const float (&xn)[SMAX],
                                               float (&r)[SMAX], float (&A)[SMAX][SMAX]) {
  const int S = p.S;
  const float sq = sqrtf(p.dt);
#pragma unroll
  for (int i = 0; i < SMAX; ++i)
#pragma unroll
    for (int j = 0; j < SMAX; ++j) A[i][j] = 0.f;
  if (p.sde_kind == VISDE_SDE_OU) {
    r[0] = xn[0] - (xt[0] + th[0] * (th[1] - xt[0]) * p.dt);
    A[0][0] = th[2] * sq;
  } else if (p.sde_kind == VISDE_SDE_LV) {
    if (SMAX >= 2) {
      float u = xt[0], v = xt[1 % SMAX];
      float f0 = th[0] * u - th[1] * u * v;
      float f1 = th[1] * u * v - th[2] * v;
      r[0] = xn[0] - (u + f0 * p.dt);
      r[1 % SMAX] = xn[1 % SMAX] - (v + f1 * p.dt);
      LvDiff d = lv_diffusion(u, v, th[0], th[1], th[2]);
      A[0][0] = d.L00 * sq;
      A[1 % SMAX][0] = d.L10 * sq;
      A[1 % SMAX][1 % SMAX] = d.L11 * sq;
    }
  } else {
    const float* f = p.drift + (b * p.T + t) * S;
    const float* D = p.diffusion + (b * p.T + t) * (int64_t)S * S;
#pragma unroll
    for (int i = 0; i < SMAX; ++i) r[i] = i < S ? xn[i] - (xt[i] + f[i] * p.dt) : 0.f;
    load_tril_scaled<SMAX>(D, S, sq, A, p.vec16 != 0);
  }
}

template <int SMAX>
__device__ __forceinline__ void gen_transition(const ElboParams& p, int64_t b, int64_t t,
                                               const float (&zt)[SMAX], const float (&zn)[SMAX],
                                               float (&r)[SMAX], float (&A)[SMAX][SMAX]) {
  const int S = p.S;
  const float* mu = p.means + (b * p.T + t) * S;
#pragma unroll
  for (int i = 0; i < SMAX; ++i) r[i] = i < S ? zn[i] - (zt[i] + mu[i] * p.dt) : 0.f;
  load_tril_scaled<SMAX>(p.chol + (b * p.T + t) * (int64_t)S * S, S, sqrtf(p.dt), A, p.vec16 != 0);
}

// observation log-likelihood at grid index tau (all observations whose idx == tau)
template <int SMAX>
__device__ __forceinline__ float obs_term(const ElboParams& p, int64_t tau, const float (&x)[SMAX],
                                          float g_obs, float (&gx)[SMAX], bool want_grad) {
  float lp = 0.f;
  const visde_obs& o = p.obs;
  for (int k = 0; k < o.n_obs; ++k) {
    if ((int64_t)o.idx[k] != tau) continue;
    for (int d = 0; d < o.obs_dim; ++d) {
      float pred = 0.f;
      if (o.obs_matrix) {
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
          if (s < p.S) pred += o.obs_matrix[d * p.S + s] * x[s];
      } else {
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
          if (s == d) pred = x[s];
      }
      float diff = o.values[k * o.obs_dim + d] - pred;
      lp += -0.5f * diff * diff / o.variance - 0.5f * logf(6.283185307179586f * o.variance);
      if (want_grad) {
        float gp = g_obs * diff / o.variance;
        if (o.obs_matrix) {
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
            if (s < p.S) gx[s] += gp * o.obs_matrix[d * p.S + s];
        } else {
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
            if (s == d) gx[s] += gp;
        }
      }
    }
  }
  return lp;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kElboMaxThreads / 32; ++i) s += (i < (int)(blockDim.x >> 5)) ? red[i] : 0.f;
  return s;
}

template <int SMAX>
__global__ void __launch_bounds__(SMAX <= 4 ? kElboMaxThreads : kElboThreads) elbo_fwd_kernel(ElboParams p) {
  __shared__ float red[kElboMaxThreads / 32];
  const int S = p.S;
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    float th[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.sde_kind != VISDE_SDE_GENERIC)
      for (int q = 0; q < 3; ++q) th[q] = p.theta[b * p.P + q];
    float s_obs = 0.f, s_sde = 0.f, s_gen = 0.f, s_jac = 0.f;
    for (int64_t tau = threadIdx.x; tau <= p.T; tau += blockDim.x) {
      float zt[SMAX], xt[SMAX];
      load_vec<SMAX>(p.z + (b * (p.T + 1) + tau) * S, S, zt);
      to_state_vec<SMAX>(S, p.pos_mask, zt, xt);
      if (tau >= 1) {
#pragma unroll
        for (int i = 0; i < SMAX; ++i)
          if (i < S && ((p.pos_mask >> i) & 1u)) s_jac += logsigmoid_f(zt[i]);
      }
      float dummy[SMAX];
      s_obs += obs_term<SMAX>(p, tau, xt, 0.f, dummy, false);
      if (tau < p.T) {
        float zn[SMAX], xn[SMAX], r[SMAX], y[SMAX], w[SMAX], A[SMAX][SMAX];
        load_vec<SMAX>(p.z + (b * (p.T + 1) + tau + 1) * S, S, zn);
        to_state_vec<SMAX>(S, p.pos_mask, zn, xn);
        gen_transition<SMAX>(p, b, tau, zt, zn, r, A);
        s_gen += gauss_solve<SMAX>(S, A, r, y, w, false);
        sde_transition<SMAX>(p, b, tau, th, xt, xn, r, A);
        s_sde += gauss_solve<SMAX>(S, A, r, y, w, false);
      }
    }
    s_obs = block_sum(s_obs, red);
    s_sde = block_sum(s_sde, red);
    s_gen = block_sum(s_gen, red);
    s_jac = block_sum(s_jac, red);
    if (threadIdx.x == 0) {
      float* o = p.terms + b * 4;
      o[0] = s_obs;
      o[1] = s_sde;
      o[2] = s_gen;
      o[3] = s_jac;
    }
  }
}

// Backward.  A thread owns grid point tau and solves ITS transition (tau -> tau + 1) once; the cotangents that transition sends
// to its end point (d lp / d z_{tau+1} = -w, d lp / d x_{tau+1} = -w) are handed to the thread of tau + 1 through shared
// memory (slot tau + 1 of the chunk; the last slot carries over to the next chunk of the trajectory).  The first version
// re-solved transition tau - 1 in thread tau: twice the triangular solves and twice the factor-block reads.
template <int SMAX>
__global__ void __launch_bounds__(SMAX <= 4 ? kElboMaxThreads : kElboThreads, SMAX <= 4 ? 1 : (SMAX <= 10 ? 3 : 2)) elbo_bwd_kernel(ElboParams p) {
  constexpr int NTMAX = SMAX <= 4 ? kElboMaxThreads : kElboThreads;
  __shared__ float red[kElboMaxThreads / 32];
  __shared__ float nxt[2][NTMAX + 1][SMAX];  // [0]: cotangent of z_next (generative term), [1]: of x_next (SDE term)
  const int S = p.S, NT = blockDim.x;
  const float sq = sqrtf(p.dt);
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    float th[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.sde_kind != VISDE_SDE_GENERIC)
      for (int q = 0; q < 3; ++q) th[q] = p.theta[b * p.P + q];
    const float g_obs = p.g_terms[b * 4 + 0], g_sde = p.g_terms[b * 4 + 1];
    const float g_gen = p.g_terms[b * 4 + 2], g_jac = p.g_terms[b * 4 + 3];
    float gth[3] = {0.f, 0.f, 0.f};
    __syncthreads();  // the previous trajectory's readers are done with the hand-over slots
    if (threadIdx.x < SMAX) nxt[0][0][threadIdx.x] = nxt[1][0][threadIdx.x] = 0.f;  // nothing arrives at tau = 0
    for (int64_t c0 = 0; c0 <= p.T; c0 += NT) {
      const int64_t tau = c0 + threadIdx.x;
      const bool live = tau <= p.T;
      float zt[SMAX], xt[SMAX], spg[SMAX], gz[SMAX], gx[SMAX], nz[SMAX], nx[SMAX];
#pragma unroll
      for (int i = 0; i < SMAX; ++i) gz[i] = gx[i] = nz[i] = nx[i] = 0.f, spg[i] = 1.f;
      if (live) {
        load_vec<SMAX>(p.z + (b * (p.T + 1) + tau) * S, S, zt);
        to_state_vec<SMAX>(S, p.pos_mask, zt, xt);
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
          const bool pos = i < S && ((p.pos_mask >> i) & 1u);
          spg[i] = pos ? softplus_grad_f(zt[i]) : 1.f;
          if (pos && tau >= 1) gz[i] += g_jac * (1.f - 1.f / (1.f + expf(-zt[i])));
        }
        obs_term<SMAX>(p, tau, xt, g_obs, gx, true);
      }
      if (live && tau < p.T) {
        float r[SMAX], y[SMAX], w[SMAX], A[SMAX][SMAX];
        const int64_t row = b * p.T + tau;
        float zn[SMAX], xn[SMAX];
        load_vec<SMAX>(p.z + (b * (p.T + 1) + tau + 1) * S, S, zn);
        to_state_vec<SMAX>(S, p.pos_mask, zn, xn);
        // generative (variational) transition: gradients w.r.t. z_t, z_{t+1}, mu_t, L_t
        gen_transition<SMAX>(p, b, tau, zt, zn, r, A);
        gauss_solve<SMAX>(S, A, r, y, w, true);
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
          if (i < S) {
            gz[i] += g_gen * w[i];
            nz[i] = -g_gen * w[i];
            p.g_means[row * S + i] = g_gen * w[i] * p.dt;
          }
        }
        store_tril_grad<SMAX>(p.g_chol + row * (int64_t)S * S, S, g_gen * sq, w, y, A, p.vec16 != 0);
        // SDE transition: gradients w.r.t. x_t (direct + through f, D), x_{t+1} and theta
        sde_transition<SMAX>(p, b, tau, th, xt, xn, r, A);
        gauss_solve<SMAX>(S, A, r, y, w, true);
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
          gx[i] += g_sde * w[i];  // through the mean's x_t term
          nx[i] = -g_sde * w[i];
        }
        if (p.sde_kind == VISDE_SDE_OU) {
          float gf = g_sde * w[0] * p.dt;
          float gD = g_sde * sq * (w[0] * y[0] - 1.f / A[0][0]);
          gx[0] += gf * (-th[0]);
          gth[0] += gf * (th[1] - xt[0]);
          gth[1] += gf * th[0];
          gth[2] += gD;
        } else if (p.sde_kind == VISDE_SDE_LV) {
          if (SMAX >= 2) {
            const float u = xt[0], v = xt[1 % SMAX], t1 = th[0], t2 = th[1], t3 = th[2], uv = u * v;
            const float w0 = w[0], w1 = w[1 % SMAX], y0 = y[0], y1 = y[1 % SMAX];
            float gf0 = g_sde * w0 * p.dt, gf1 = g_sde * w1 * p.dt;
            float gL00 = g_sde * sq * (w0 * y0 - 1.f / A[0][0]);
            float gL10 = g_sde * sq * (w1 * y0);
            float gL11 = g_sde * sq * (w1 * y1 - 1.f / A[1 % SMAX][1 % SMAX]);
            LvDiff d = lv_diffusion(u, v, t1, t2, t3);
            // reverse through the Cholesky with torch.clamp's inclusive pass-through
            float g_e = (d.e >= 1e-6f) ? gL11 * 0.5f / d.L11 : 0.f;
            float g_b22 = g_e;
            float g_L10 = gL10 - 2.f * d.L10 * g_e;
            float g_b12 = g_L10 / d.d00;
            float g_d00 = -g_L10 * d.b12 / (d.d00 * d.d00);
            float g_L00 = gL00 + ((d.L00 >= 1e-6f) ? g_d00 : 0.f);
            float g_b11 = (d.b11 >= 1e-6f) ? g_L00 * 0.5f / d.L00 : 0.f;
            float g_t1 = g_b11 * u + gf0 * u;
            float g_t2 = (g_b11 - g_b12 + g_b22) * uv + (gf1 - gf0) * uv;
            float g_t3 = g_b22 * v - gf1 * v;
            float g_uv = (g_b11 - g_b12 + g_b22) * t2 + (gf1 - gf0) * t2;
            float g_u = g_b11 * t1 + gf0 * t1 + g_uv * v;
            float g_v = g_b22 * t3 - gf1 * t3 + g_uv * u;
            gx[0] += g_u;
            gx[1 % SMAX] += g_v;
            gth[0] += g_t1;
            gth[1] += g_t2;
            gth[2] += g_t3;
          }
        } else {
#pragma unroll
          for (int i = 0; i < SMAX; ++i)
            if (i < S) p.g_drift[row * S + i] = g_sde * w[i] * p.dt;
          store_tril_grad<SMAX>(p.g_diffusion + row * (int64_t)S * S, S, g_sde * sq, w, y, A, p.vec16 != 0);
        }
      }
      // hand the end-point cotangents to the thread of tau + 1
#pragma unroll
      for (int i = 0; i < SMAX; ++i) {
        nxt[0][threadIdx.x + 1][i] = nz[i];
        nxt[1][threadIdx.x + 1][i] = nx[i];
      }
      __syncthreads();
      if (live) {
#pragma unroll
        for (int i = 0; i < SMAX; ++i)
          if (i < S) p.g_z[(b * (p.T + 1) + tau) * S + i] = (gz[i] + nxt[0][threadIdx.x][i]) + (gx[i] + nxt[1][threadIdx.x][i]) * spg[i];
      }
      if (threadIdx.x == 0) {  // carry: the last thread's hand-over belongs to the first grid point of the next chunk
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
          nxt[0][0][i] = nxt[0][NT][i];
          nxt[1][0][i] = nxt[1][NT][i];
        }
      }
      __syncthreads();
    }
    float t0 = block_sum(gth[0], red), t1 = block_sum(gth[1], red), t2 = block_sum(gth[2], red);
    if (threadIdx.x == 0) {
      for (int q = 0; q < p.P; ++q) p.g_theta[b * p.P + q] = 0.f;
      if (p.sde_kind != VISDE_SDE_GENERIC) {
        p.g_theta[b * p.P + 0] = t0;
        p.g_theta[b * p.P + 1] = t1;
        p.g_theta[b * p.P + 2] = t2;
      }
    }
  }
}

int elbo_grid(int64_t B) {
  int64_t cap = 148 * 16;
  return (int)(B < cap ? B : cap);
}

// one block per trajectory: with few trajectories every extra pass over tau is a serial HBM round trip
int elbo_threads(int64_t B, int64_t T, int smax) {
  if (B > 148 * 4 || smax > 4) return kElboThreads;  // wide-state instantiations keep the 128-thread register budget
  int64_t t = (T + 1 + 31) / 32 * 32;
  return (int)(t < kElboThreads ? kElboThreads : (t > kElboMaxThreads ? kElboMaxThreads : t));
}


template <int SMAX>
int launch_both(const ElboParams& p, cudaStream_t st, bool bwd) {
  if (bwd)
    elbo_bwd_kernel<SMAX><<<elbo_grid(p.B), elbo_threads(p.B, p.T, SMAX), 0, st>>>(p);
  else
    elbo_fwd_kernel<SMAX><<<elbo_grid(p.B), elbo_threads(p.B, p.T, SMAX), 0, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int dispatch(const ElboParams& p_in, cudaStream_t st, bool bwd) {
  if (p_in.B == 0) return VISDE_OK;
  ElboParams p = p_in;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };  // NULL counts as aligned
  p.vec16 = al(p.chol) && al(p.diffusion) && al(p.g_chol) && al(p.g_diffusion) && (p.S * p.S) % 4 == 0;
  if (p.S <= 1) return launch_both<1>(p, st, bwd);
  if (p.S <= 2) return launch_both<2>(p, st, bwd);
  if (p.S <= 4) return launch_both<4>(p, st, bwd);
  if (p.S <= 6) return launch_both<6>(p, st, bwd);
  if (p.S <= 8) return launch_both<8>(p, st, bwd);
  if (p.S <= 10) return launch_both<10>(p, st, bwd);  // BASELINE config 5 (10-D Lorenz-96): 100 instead of 256 matrix registers
  if (p.S <= 12) return launch_both<12>(p, st, bwd);
  return launch_both<16>(p, st, bwd);
}

}  // namespace

int launch_elbo_fwd(const ElboParams& p, cudaStream_t st) { return dispatch(p, st, false); }
int launch_elbo_bwd(const ElboParams& p, cudaStream_t st) { return dispatch(p, st, true); }

}  // namespace visde
