// Euler-Maruyama prior simulator (forward + reverse-mode) for the built-in SDE functors.
//
// Restates core/euler_maruyama.py:11-45 as used by the theta pre-training loop
// (inference/trainer.py:208-259: 4096 simulations x 1000 Adam steps, gradient of an MSE at the observation
// grid points with respect to theta):
//     x_{t+1} = x_t + f(x_t, th) dt + D(x_t, th) eps_t sqrt(dt);   positive dims clamped to >= 1e-6   (:37-42)
// with f, D the Ornstein-Uhlenbeck / Lotka-Volterra functors of examples/*.py (same algebra as elbo.cu).
// User SDEs are stepped in PyTorch by the host mirror (viforsdes_b200/euler_maruyama.py), as BASELINE.json asks.
//
// Layout: one LANE per trajectory (state, theta and the theta-adjoint stay in registers for all T steps), 32 trajectories
// per CTA so that 4096 trajectories spread over 128 SMs; a second warp per CTA moves the operands (see em_fwd_kernel).  The per-step operands of a trajectory are 4-8 bytes at a
// stride of T*S floats, so the warp moves them in chunks of 32/S steps through a padded shared-memory tile: every
// global access is one contiguous <=128-byte row segment of one trajectory, 32 independent rows in flight.
// noise == NULL draws eps in the kernel: Philox4x32-10 keyed by `seed`, counter (t, b), Box-Muller; the backward
// regenerates the same draws, so the [B,T,S] noise tensor is never materialised.
#include "common.cuh"

namespace visde {
namespace {

constexpr float kEmClamp = 1e-6f;  // core/euler_maruyama.py:42
constexpr int kTileLd = 33;        // padded row length of the warp tiles (bank-conflict free both ways)

// ---- Philox4x32-10 (Salmon et al. 2011; constants as published) ------------------------------------------------
struct U4 { uint32_t x, y, z, w; };
__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
// four standard normals for (trajectory b, step t): u = (top 24 bits + 0.5) 2^-24 in (0, 1); Box-Muller pairs
// `group` selects the block of four for state dims 4 * group .. 4 * group + 3 (wide state spaces; group 0 = the original stream)
__device__ __forceinline__ void philox_normal4(uint64_t seed, int64_t b, int64_t t, float (&n)[4], uint32_t group = 0) {
  const U4 r = philox4x32_10(U4{(uint32_t)t, (uint32_t)((uint64_t)t >> 32) | (group << 28), (uint32_t)b, (uint32_t)((uint64_t)b >> 32)},
                             (uint32_t)seed, (uint32_t)(seed >> 32));
  const float s = 5.9604644775390625e-8f;  // 2^-24
  const float u0 = ((float)(r.x >> 8) + 0.5f) * s, u1 = ((float)(r.y >> 8) + 0.5f) * s;
  const float u2 = ((float)(r.z >> 8) + 0.5f) * s, u3 = ((float)(r.w >> 8) + 0.5f) * s;
  // accurate log (u near 1 needs relative accuracy); fast sin / cos: |error| < 4e-7 in [-pi, pi], i.e. ~2e-6 on a draw
  const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u1 - 3.14159265358979f, &sa, &ca);  // argument in [-pi, pi): the intrinsic's accurate range
  __sincosf(6.283185307179586f * u3 - 3.14159265358979f, &sb, &cb);
  sa = -sa; ca = -ca; sb = -sb; cb = -cb;
  n[0] = ra * ca;
  n[1] = ra * sa;
  n[2] = rb * cb;
  n[3] = rb * sb;
}

// ---- SDE functors: step() and its reverse-mode -------------------------------------------------------------------
// step: xn = x + f dt + D eps sqrt_dt (before the clamp).  step_bwd: given a = d loss / d xn, accumulates gth and
// returns d loss / d x (through this transition only).
struct OuModel {
  static constexpr int S = 1, P = 3;
  __device__ static void step(const float (&x)[S], const float (&th)[P], const float (&e)[S], float dt, float sq,
                              float (&xn)[S]) {
    xn[0] = x[0] + th[0] * (th[1] - x[0]) * dt + th[2] * e[0] * sq;  // examples/ornstein_uhlenbeck.py:18-30
  }
  __device__ static void step_bwd(const float (&x)[S], const float (&th)[P], const float (&e)[S], float dt, float sq,
                                  const float (&a)[S], float (&gth)[P], float (&gx)[S]) {
    const float gf = a[0] * dt;
    gth[0] = fmaf(gf, th[1] - x[0], gth[0]);
    gth[1] = fmaf(gf, th[0], gth[1]);
    gth[2] = fmaf(a[0] * sq, e[0], gth[2]);
    gx[0] = a[0] - gf * th[0];
  }
};

struct LvModel {
  static constexpr int S = 2, P = 3;
  // examples/lotka_volterra.py:18-46 (drift, Cholesky factor of the diffusion matrix with its three clamps)
  // Cholesky factor through MUFU.RSQ (L = m rsqrt(m), 1 / L = rsqrt(m); ~2 ulp) instead of IEEE sqrt + divisions: the
  // transition is the serial chain of the kernel.  L00 >= sqrt(1e-6) = 1e-3, so the reference's clamp(L00, 1e-6) never binds.
  struct Chol {
    float b11, b12, b22, r00, L00, L10, ee, r11, L11;
  };
  __device__ static Chol chol(float u, float v, float uv, const float (&th)[P]) {
    Chol c;
    c.b11 = th[0] * u + th[1] * uv;
    c.b12 = -th[1] * uv;
    c.b22 = th[2] * v + th[1] * uv;
    const float m11 = fmaxf(c.b11, 1e-6f);
    c.r00 = rsqrtf(m11);
    c.L00 = m11 * c.r00;
    c.L10 = c.b12 * c.r00;
    c.ee = c.b22 - c.L10 * c.L10;
    const float m22 = fmaxf(c.ee, 1e-6f);
    c.r11 = rsqrtf(m22);
    c.L11 = m22 * c.r11;
    return c;
  }
  __device__ static void step(const float (&x)[S], const float (&th)[P], const float (&e)[S], float dt, float sq,
                              float (&xn)[S]) {
    const float u = x[0], v = x[1], uv = u * v;
    const float f0 = th[0] * u - th[1] * uv, f1 = th[1] * uv - th[2] * v;
    const Chol c = chol(u, v, uv, th);
    xn[0] = u + f0 * dt + (c.L00 * e[0]) * sq;
    xn[1] = v + f1 * dt + (c.L10 * e[0] + c.L11 * e[1]) * sq;
  }
  __device__ static void step_bwd(const float (&x)[S], const float (&th)[P], const float (&e)[S], float dt, float sq,
                                  const float (&a)[S], float (&gth)[P], float (&gx)[S]) {
    const float u = x[0], v = x[1], uv = u * v;
    const Chol c = chol(u, v, uv, th);
    // noise term
    float gL00 = a[0] * e[0] * sq, gL10 = a[1] * e[0] * sq;
    const float gL11 = a[1] * e[1] * sq;
    const float ge = c.ee >= 1e-6f ? 0.5f * gL11 * c.r11 : 0.f;  // torch clamp(min): gradient passes where input >= min
    const float gb22 = ge;
    gL10 -= 2.f * c.L10 * ge;
    const float gb12 = gL10 * c.r00;
    gL00 -= gL10 * c.L10 * c.r00;  // through the denominator of L10 = b12 / L00
    const float gb11 = c.b11 >= 1e-6f ? 0.5f * gL00 * c.r00 : 0.f;
    // drift
    const float gf0 = a[0] * dt, gf1 = a[1] * dt;
    const float g1 = gf0 + gb11;                         // d / d(th0 u)
    const float g3 = gb22 - gf1;                         // d / d(th2 v)
    const float gq = gf1 - gf0 + gb11 - gb12 + gb22;     // d / d(th1 u v)
    gth[0] = fmaf(g1, u, gth[0]);
    gth[1] = fmaf(gq, uv, gth[1]);
    gth[2] = fmaf(g3, v, gth[2]);
    gx[0] = a[0] + th[0] * g1 + th[1] * v * gq;
    gx[1] = a[1] + th[2] * g3 + th[1] * u * gq;
  }
};

// ---- warp tile <-> global rows ---------------------------------------------------------------------------------
// tile[r][c] <- row r of 32 trajectories, `ncols` contiguous floats starting at base + r * row_stride
__device__ __forceinline__ void tile_load(float* tile, const float* __restrict__ base, int64_t row_stride, int nrows, int ncols, int lane) {
  // two passes of 16 rows, every load of a pass issued before its first store: two memory latencies per tile
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += 16) {
    float v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = (r0 + r < nrows && lane < ncols) ? base[(r0 + r) * row_stride + lane] : 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r0 + r < nrows && lane < ncols) tile[(r0 + r) * kTileLd + lane] = v[r];
  }
}
__device__ __forceinline__ void tile_store(const float* tile, float* base, int64_t row_stride, int nrows, int ncols, int lane) {
#pragma unroll 8
  for (int r = 0; r < 32; ++r)
    if (r < nrows && lane < ncols) base[r * row_stride + lane] = tile[r * kTileLd + lane];
}

struct EmParams {
  int64_t B, T;
  uint32_t pos_mask;
  float dt, sqrt_dt;
  uint64_t seed;
  const float* x0;
  const float* theta;
  const float* noise;    // [B,T,S] or nullptr (Philox)
  float* paths;          // fwd out / bwd in [B,T+1,S]
  const float* g_paths;  // bwd in [B,T+1,S]
  float* grad_x0;        // bwd out [B,S] or nullptr
  float* grad_theta;     // bwd out [B,P]
};

// noise of chunk [t0, t0 + tc) for the 32 trajectories of this CTA -> tile (injected: coalesced row loads; Philox: each lane
// generates its own trajectory's draws, the tc blocks are independent so the integer rounds of several steps overlap)
template <int S>
__device__ __forceinline__ void chunk_noise(const EmParams& p, float* tile, int64_t b0, int nrows, int64_t t0, int tc, int lane) {
  if (p.noise) {
    tile_load(tile, p.noise + (b0 * p.T + t0) * S, p.T * S, nrows, tc * S, lane);
  } else {
#pragma unroll 4
    for (int j = 0; j < tc; ++j) {
      float n4[4];
      philox_normal4(p.seed, b0 + lane, t0 + j, n4);
#pragma unroll
      for (int s = 0; s < S; ++s) tile[lane * kTileLd + j * S + s] = n4[s];
    }
  }
}

// Four warps per CTA.  Warp 0 is the stepper: state in registers, operands and results only through shared-memory tiles, so
// nothing on its serial chain ever waits on HBM or on the Philox rounds.  The others are movers: while warp 0 steps chunk c,
// warp 1 prepares the noise tile of chunk c+1 (draws or loads it) and warp 2 stores the path tile of chunk c-1 (backward:
// warps 1-3 load the x, cotangent and noise tiles of the next chunk).  One barrier per chunk.
constexpr int kEmThreads = 128;
template <class M>
__global__ void __launch_bounds__(kEmThreads) em_fwd_kernel(EmParams p) {
  constexpr int S = M::S, P = M::P, TC = 32 / S;
  __shared__ float tin[2][32 * kTileLd], tout[2][32 * kTileLd];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * 32, b = b0 + lane;
  const int nrows = (int)(p.B - b0 < 32 ? p.B - b0 : 32);
  const bool ok = lane < nrows;
  const int64_t nchunk = (p.T + TC - 1) / TC;
  float x[S], th[P];
#pragma unroll
  for (int s = 0; s < S; ++s) x[s] = ok ? p.x0[b * S + s] : 1.f;
#pragma unroll
  for (int q = 0; q < P; ++q) th[q] = ok ? p.theta[b * P + q] : 1.f;
  if (warp == 0 && ok) {
#pragma unroll
    for (int s = 0; s < S; ++s) p.paths[b * (p.T + 1) * S + s] = x[s];
  }
  if (warp == 1 && nchunk > 0) chunk_noise<S>(p, tin[0], b0, nrows, 0, (int)(p.T < TC ? p.T : TC), lane);
  __syncthreads();
  for (int64_t c = 0; c < nchunk; ++c) {
    const int64_t t0 = c * TC;
    const int tc = (int)(p.T - t0 < TC ? p.T - t0 : TC);
    if (warp == 1) {
      if (c + 1 < nchunk) {
        const int64_t t1 = t0 + TC;
        chunk_noise<S>(p, tin[(c + 1) & 1], b0, nrows, t1, (int)(p.T - t1 < TC ? p.T - t1 : TC), lane);
      }
    } else if (warp == 2) {
      if (c > 0) tile_store(tout[(c - 1) & 1], p.paths + (b0 * (p.T + 1) + t0 - TC + 1) * S, (p.T + 1) * S, nrows, TC * S, lane);
    } else if (warp == 0) {
      const float* ti = tin[c & 1];
      float* to = tout[c & 1];
      for (int j = 0; j < tc; ++j) {
        float e[S], xn[S];
#pragma unroll
        for (int s = 0; s < S; ++s) e[s] = ti[lane * kTileLd + j * S + s];
        M::step(x, th, e, p.dt, p.sqrt_dt, xn);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          x[s] = ((p.pos_mask >> s) & 1u) ? fmaxf(xn[s], kEmClamp) : xn[s];
          to[lane * kTileLd + j * S + s] = x[s];
        }
      }
    }
    __syncthreads();
  }
  if (warp == 2 && nchunk > 0) {
    const int64_t t0 = (nchunk - 1) * TC;
    tile_store(tout[(nchunk - 1) & 1], p.paths + (b0 * (p.T + 1) + t0 + 1) * S, (p.T + 1) * S, nrows, (int)(p.T - t0) * S, lane);
  }
}

template <int S>
__device__ __forceinline__ void bwd_chunk_operands(const EmParams& p, float* tx, float* tg, float* tn, int64_t b0, int nrows,
                                                   int64_t t0, int tc, int lane, int warp) {
  // x_t for t in the chunk (warp 1), cotangents of x_{t+1} (warp 2), and the noise of the chunk (warp 3)
  if (warp == 1) tile_load(tx, p.paths + (b0 * (p.T + 1) + t0) * S, (p.T + 1) * S, nrows, tc * S, lane);
  if (warp == 2) tile_load(tg, p.g_paths + (b0 * (p.T + 1) + t0 + 1) * S, (p.T + 1) * S, nrows, tc * S, lane);
  if (warp == 3) chunk_noise<S>(p, tn, b0, nrows, t0, tc, lane);
}

template <class M>
__global__ void __launch_bounds__(kEmThreads) em_bwd_kernel(EmParams p) {
  constexpr int S = M::S, P = M::P, TC = 32 / S;
  __shared__ float tn[2][32 * kTileLd], tx[2][32 * kTileLd], tg[2][32 * kTileLd];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * 32, b = b0 + lane;
  const int nrows = (int)(p.B - b0 < 32 ? p.B - b0 : 32);
  const bool ok = lane < nrows;
  float th[P], gth[P], a[S];
#pragma unroll
  for (int q = 0; q < P; ++q) {
    th[q] = ok ? p.theta[b * P + q] : 1.f;
    gth[q] = 0.f;
  }
#pragma unroll
  for (int s = 0; s < S; ++s) a[s] = 0.f;
  const int64_t nchunk = (p.T + TC - 1) / TC;
  auto chunk_len = [&](int64_t c) { return (int)(p.T - c * TC < TC ? p.T - c * TC : TC); };
  if (warp >= 1 && nchunk > 0) {
    const int64_t c = nchunk - 1;
    bwd_chunk_operands<S>(p, tx[c & 1], tg[c & 1], tn[c & 1], b0, nrows, c * TC, chunk_len(c), lane, warp);
  }
  __syncthreads();
  for (int64_t c = nchunk - 1; c >= 0; --c) {
    if (warp >= 1) {
      if (c > 0) bwd_chunk_operands<S>(p, tx[(c - 1) & 1], tg[(c - 1) & 1], tn[(c - 1) & 1], b0, nrows, (c - 1) * TC, TC, lane, warp);
    } else {
      const float *cx = tx[c & 1], *cg = tg[c & 1], *cn = tn[c & 1];
      for (int j = chunk_len(c) - 1; j >= 0; --j) {
        float x[S], e[S], xn[S], gx[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          x[s] = ok ? cx[lane * kTileLd + j * S + s] : 1.f;
          a[s] += ok ? cg[lane * kTileLd + j * S + s] : 0.f;
          e[s] = ok ? cn[lane * kTileLd + j * S + s] : 0.f;
        }
        // clamp(min=1e-6) passes the gradient where the unclamped value is >= 1e-6: recompute it exactly as the forward did
        M::step(x, th, e, p.dt, p.sqrt_dt, xn);
#pragma unroll
        for (int s = 0; s < S; ++s)
          if (((p.pos_mask >> s) & 1u) && !(xn[s] >= kEmClamp)) a[s] = 0.f;
        M::step_bwd(x, th, e, p.dt, p.sqrt_dt, a, gth, gx);
#pragma unroll
        for (int s = 0; s < S; ++s) a[s] = gx[s];
      }
    }
    __syncthreads();
  }
  if (warp == 0 && ok) {
#pragma unroll
    for (int q = 0; q < P; ++q) p.grad_theta[b * P + q] = gth[q];
    if (p.grad_x0) {
#pragma unroll
      for (int s = 0; s < S; ++s) p.grad_x0[b * S + s] = a[s] + p.g_paths[b * (p.T + 1) * S + s];
    }
  }
}

__global__ void philox_normal_kernel(uint64_t seed, int64_t B, int64_t T, int S, float* out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B * T) return;
  for (int g = 0; 4 * g < S; ++g) {
    float n4[4];
    philox_normal4(seed, k / T, k % T, n4, (uint32_t)g);
    for (int s = 4 * g; s < S && s < 4 * g + 4; ++s) out[k * S + s] = n4[s - 4 * g];
  }
}

int em_model_dims(int sde_kind, int* S, int* P) {
  if (sde_kind == VISDE_SDE_OU) { *S = OuModel::S; *P = OuModel::P; return VISDE_OK; }
  if (sde_kind == VISDE_SDE_LV) { *S = LvModel::S; *P = LvModel::P; return VISDE_OK; }
  set_error("euler_maruyama: sde_kind %d has no device functor (user SDEs are stepped in PyTorch by the host mirror)", sde_kind);
  return VISDE_EINVAL;
}

}  // namespace
}  // namespace visde

using namespace visde;

extern "C" {

int visde_em_fwd(int64_t B, int64_t T, int sde_kind, uint32_t positive_mask, float dt, const float* x0,
                 const float* theta, const float* noise, uint64_t seed, float* paths, void* stream) {
  int S, P;
  int rc = em_model_dims(sde_kind, &S, &P);
  if (rc) return rc;
  VISDE_REQUIRE(dt > 0.f, "dt must be positive, got %g", (double)dt);  // core/euler_maruyama.py:20-21
  VISDE_REQUIRE(B >= 0 && T >= 0, "euler_maruyama: negative batch / step count");
  if (B == 0) return VISDE_OK;
  VISDE_REQUIRE(x0 && theta && paths, "euler_maruyama: NULL tensor argument");
  EmParams p{};
  p.B = B; p.T = T; p.pos_mask = positive_mask; p.dt = dt; p.sqrt_dt = (float)sqrt((double)dt); p.seed = seed;
  p.x0 = x0; p.theta = theta; p.noise = noise; p.paths = paths;
  const unsigned grid = (unsigned)((B + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (sde_kind == VISDE_SDE_OU) em_fwd_kernel<OuModel><<<grid, kEmThreads, 0, st>>>(p);
  else em_fwd_kernel<LvModel><<<grid, kEmThreads, 0, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int visde_em_bwd(int64_t B, int64_t T, int sde_kind, uint32_t positive_mask, float dt, const float* theta,
                 const float* noise, uint64_t seed, const float* paths, const float* g_paths, float* grad_x0,
                 float* grad_theta, void* stream) {
  int S, P;
  int rc = em_model_dims(sde_kind, &S, &P);
  if (rc) return rc;
  VISDE_REQUIRE(dt > 0.f, "dt must be positive, got %g", (double)dt);
  VISDE_REQUIRE(B >= 0 && T >= 0, "euler_maruyama: negative batch / step count");
  if (B == 0) return VISDE_OK;
  VISDE_REQUIRE(theta && paths && g_paths && grad_theta, "euler_maruyama backward: NULL tensor argument");
  EmParams p{};
  p.B = B; p.T = T; p.pos_mask = positive_mask; p.dt = dt; p.sqrt_dt = (float)sqrt((double)dt); p.seed = seed;
  p.theta = theta; p.noise = noise; p.paths = const_cast<float*>(paths); p.g_paths = g_paths;
  p.grad_x0 = grad_x0; p.grad_theta = grad_theta;
  const unsigned grid = (unsigned)((B + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (sde_kind == VISDE_SDE_OU) em_bwd_kernel<OuModel><<<grid, kEmThreads, 0, st>>>(p);
  else em_bwd_kernel<LvModel><<<grid, kEmThreads, 0, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int visde_philox_normal(uint64_t seed, int64_t B, int64_t T, int32_t S, float* out, void* stream) {
  VISDE_REQUIRE(S >= 1 && S <= VISDE_MAX_STATE, "philox_normal: 1 <= S <= %d (one Philox block per four state dims), got %d",
                VISDE_MAX_STATE, S);
  VISDE_REQUIRE(B >= 0 && T >= 0, "philox_normal: negative size");
  if (B * T == 0) return VISDE_OK;
  VISDE_REQUIRE(out, "philox_normal: out is NULL");
  philox_normal_kernel<<<(unsigned)((B * T + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, B, T, S, out);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // extern "C"
