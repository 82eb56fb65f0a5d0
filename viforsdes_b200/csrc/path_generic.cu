// Generic (any H <= 256, NL <= 4, S <= 16) path-sampling recurrence kernels.
//
// One CTA walks one trajectory through all T steps; weights are streamed from global memory
// (L2-resident) every step.  This family exists for shape coverage and as an independent
// on-device cross-check of the register-resident family in path_fast.cu; it follows the step
// math of models/head.py:68-97 and the BPTT of kernels/backward.py:257-624 (re-derived, not
// translated: gate pre-activation gradients are emitted to HBM and every weight gradient is
// formed afterwards by time-parallel GEMMs instead of per-step global atomics).
#include "common.cuh"

namespace visde {

namespace {

constexpr int kThreads = 256;

struct GenericSmem {
  float* h;     // [NL][H]
  float* gi;    // [3H]
  float* gh;    // [3H]
  float* gth;   // [3H]   theta rows of W_ih_l0 applied to theta_b
  float* out;   // [n_out]
  float* z;     // [S]
  float* eps;   // [S]
  __device__ GenericSmem(float* base, int NL, int H, int n_out, int S) {
    h = base;
    gi = h + NL * H;
    gh = gi + 3 * H;
    gth = gh + 3 * H;
    out = gth + 3 * H;
    z = out + n_out;
    eps = z + S;
  }
  static size_t bytes(int NL, int H, int n_out, int S) {
    return sizeof(float) * (size_t)(NL * H + 9 * H + n_out + 2 * S);
  }
};

__global__ void __launch_bounds__(kThreads) path_fwd_generic_kernel(PathParams p) {
  extern __shared__ float smem_f[];
  const int H = p.H, S = p.S, NL = p.NL, G = 3 * p.H;
  const int ld0 = p.S + p.C + p.P;
  GenericSmem sm(smem_f, NL, H, p.n_out, S);
  const int tid = threadIdx.x;

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    for (int i = tid; i < NL * H; i += kThreads) sm.h[i] = 0.f;
    for (int j = tid; j < G; j += kThreads) {
      float acc = 0.f;
      for (int q = 0; q < p.P; ++q) acc += p.w_ih[0][(int64_t)j * ld0 + S + p.C + q] * p.theta[b * p.P + q];
      sm.gth[j] = acc;
    }
    if (tid < S) {
      float v = p.x0[b * S + tid];
      sm.z[tid] = v;
      p.paths[b * (p.T + 1) * S + tid] = v;
    }
    __syncthreads();

    for (int64_t t = 0; t < p.T; ++t) {
      const int64_t row = b * p.T + t;
      for (int k = 0; k < NL; ++k) {
        // gate pre-activations, one row per thread
        for (int j = tid; j < G; j += kThreads) {
          float gi, gh = p.b_hh[k][j];
          const float* whh = p.w_hh[k] + (int64_t)j * H;
          const float* hk = sm.h + k * H;
          for (int q = 0; q < H; ++q) gh += whh[q] * hk[q];
          if (k == 0) {
            gi = p.gi_ctx[row * G + j] + sm.gth[j];
            const float* wz = p.w_ih[0] + (int64_t)j * ld0;
            for (int s = 0; s < S; ++s) gi += wz[s] * sm.z[s];
          } else {
            gi = p.b_ih[k][j];
            const float* wih = p.w_ih[k] + (int64_t)j * H;
            const float* hb = sm.h + (k - 1) * H;
            for (int q = 0; q < H; ++q) gi += wih[q] * hb[q];
          }
          sm.gi[j] = gi;
          sm.gh[j] = gh;
        }
        __syncthreads();
        for (int i = tid; i < H; i += kThreads) {
          float r = sigmoid_f(sm.gi[i] + sm.gh[i]);
          float u = sigmoid_f(sm.gi[H + i] + sm.gh[H + i]);
          float nhh = sm.gh[2 * H + i];
          float n = tanh_f(sm.gi[2 * H + i] + r * nhh);
          float hn = (1.f - u) * n + u * sm.h[k * H + i];
          sm.h[k * H + i] = hn;
          if (p.stash) {
            float* st = p.stash + (row * NL + k) * (int64_t)(kStashSlots * H);
            st[kStashR * H + i] = r;
            st[kStashU * H + i] = u;
            st[kStashN * H + i] = n;
            st[kStashNhh * H + i] = nhh;
            st[kStashH * H + i] = hn;
          }
        }
        __syncthreads();
      }
      // output projection
      for (int m = tid; m < p.n_out; m += kThreads) {
        float acc = p.out_b[m];
        const float* wo = p.out_w + (int64_t)m * H;
        const float* ht = sm.h + (NL - 1) * H;
        for (int q = 0; q < H; ++q) acc += wo[q] * ht[q];
        sm.out[m] = acc;
      }
      if (tid < S) sm.eps[tid] = p.eps[row * S + tid];
      __syncthreads();
      // reparameterised Euler-Maruyama update, one state dim per thread
      if (tid < S) {
        const int s = tid;
        float mu = sm.out[s];
        float acc = 0.f;
        float* Lrow = p.chol + (row * S + s) * S;
        for (int j = 0; j < S; ++j) {
          float L = 0.f;
          if (j <= s) {
            const int ti = s * (s + 1) / 2 + j;
            float raw = sm.out[S + ti];
            L = (j == s) ? fmaxf(raw, VISDE_DIAG_MIN) : raw;
            acc += L * sm.eps[j];
            if (p.raw) p.raw[row * p.n_tril + ti] = raw;
          }
          Lrow[j] = L;
        }
        float zn = sm.z[s] + mu * p.dt + acc * p.sqrt_dt;
        p.means[row * S + s] = mu;
        p.paths[(b * (p.T + 1) + t + 1) * S + s] = zn;
        sm.z[s] = zn;
      }
      __syncthreads();
    }
  }
}

struct GenericBwdSmem {
  float* dhc;   // [NL][H] gradient carried to h_k(t-1) through the recurrent path
  float* dh;    // [H]     total gradient of the current layer's h(t)
  float* dg;    // [4][H]
  float* sdgi;  // [3H]
  float* dout;  // [n_out]
  float* dz;    // [S]
  float* eps;   // [S]
  __device__ GenericBwdSmem(float* base, int NL, int H, int n_out, int S) {
    dhc = base;
    dh = dhc + NL * H;
    dg = dh + H;
    sdgi = dg + 4 * H;
    dout = sdgi + 3 * H;
    dz = dout + n_out;
    eps = dz + S;
  }
  static size_t bytes(int NL, int H, int n_out, int S) {
    return sizeof(float) * (size_t)(NL * H + 8 * H + n_out + 2 * S);
  }
};

__global__ void __launch_bounds__(kThreads) path_bwd_generic_kernel(PathParams p) {
  extern __shared__ float smem_f[];
  const int H = p.H, S = p.S, NL = p.NL, G = 3 * p.H;
  const int ld0 = p.S + p.C + p.P;
  GenericBwdSmem sm(smem_f, NL, H, p.n_out, S);
  const int tid = threadIdx.x;
  const int64_t srow = stash_row_floats(NL, H);

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    for (int i = tid; i < NL * H; i += kThreads) sm.dhc[i] = 0.f;
    for (int j = tid; j < G; j += kThreads) sm.sdgi[j] = 0.f;
    if (tid < S) sm.dz[tid] = 0.f;
    __syncthreads();

    for (int64_t t = p.T - 1; t >= 0; --t) {
      const int64_t row = b * p.T + t;
      if (tid < S) {
        sm.dz[tid] += p.g_paths[(b * (p.T + 1) + t + 1) * S + tid];
        sm.eps[tid] = p.eps[row * S + tid];
      }
      __syncthreads();
      // cotangent of the output projection: (d mu, d raw tril)
      for (int m = tid; m < p.n_out; m += kThreads) {
        float d;
        if (m < S) {
          d = sm.dz[m] * p.dt + p.g_means[row * S + m];
        } else {
          const int ti = m - S;
          int r = 0;
          while ((r + 1) * (r + 2) / 2 <= ti) ++r;
          const int c = ti - r * (r + 1) / 2;
          d = sm.dz[r] * sm.eps[c] * p.sqrt_dt + p.g_chol[(row * S + r) * S + c];
          if (r == c) {
            // primitives/bounds.py:20: pass iff raw >= bound or grad < 0
            float raw = p.raw[row * p.n_tril + ti];
            if (!(raw >= VISDE_DIAG_MIN || d < 0.f)) d = 0.f;
          }
        }
        sm.dout[m] = d;
        p.dout[row * p.n_out + m] = d;
      }
      __syncthreads();
      for (int i = tid; i < H; i += kThreads) {
        float acc = sm.dhc[(NL - 1) * H + i];
        for (int m = 0; m < p.n_out; ++m) acc += p.out_w[(int64_t)m * H + i] * sm.dout[m];
        sm.dh[i] = acc;
      }
      for (int k = NL - 1; k >= 0; --k) {
        for (int i = tid; i < H; i += kThreads) {
          const float* st = p.stash + (row * NL + k) * (int64_t)(kStashSlots * H);
          float r = st[kStashR * H + i], u = st[kStashU * H + i], n = st[kStashN * H + i];
          float nhh = st[kStashNhh * H + i];
          float hprev = (t > 0) ? (st - srow)[kStashH * H + i] : 0.f;
          float d = sm.dh[i];
          float dnp = d * (1.f - u) * (1.f - n * n);
          float dup = d * (hprev - n) * u * (1.f - u);
          float drp = dnp * nhh * r * (1.f - r);
          float dnh = dnp * r;
          sm.dg[0 * H + i] = drp;
          sm.dg[1 * H + i] = dup;
          sm.dg[2 * H + i] = dnp;
          sm.dg[3 * H + i] = dnh;
          sm.dhc[k * H + i] = d * u;
          float* dgo = p.dg + (row * NL + k) * (int64_t)(kDgSlots * H);
          dgo[0 * H + i] = drp;
          dgo[1 * H + i] = dup;
          dgo[2 * H + i] = dnp;
          dgo[3 * H + i] = dnh;
        }
        __syncthreads();
        for (int i = tid; i < H; i += kThreads) {
          const float* whh = p.w_hh[k];
          float acc = 0.f;
          for (int j = 0; j < H; ++j) {
            acc += whh[(int64_t)(j)*H + i] * sm.dg[j];
            acc += whh[(int64_t)(H + j) * H + i] * sm.dg[H + j];
            acc += whh[(int64_t)(2 * H + j) * H + i] * sm.dg[3 * H + j];
          }
          sm.dhc[k * H + i] += acc;
          if (k > 0) {
            const float* wih = p.w_ih[k];
            float accb = sm.dhc[(k - 1) * H + i];
            for (int j = 0; j < G; ++j) accb += wih[(int64_t)j * H + i] * sm.dg[j];
            sm.dh[i] = accb;
          }
        }
        if (k == 0) {
          if (tid < S) {
            float acc = 0.f;
            for (int j = 0; j < G; ++j) acc += p.w_ih[0][(int64_t)j * ld0 + tid] * sm.dg[j];
            sm.dz[tid] += acc;
          }
          for (int j = tid; j < G; j += kThreads) sm.sdgi[j] += sm.dg[j];
        }
        __syncthreads();
      }
    }
    if (tid < S) p.grad_x0[b * S + tid] = sm.dz[tid] + p.g_paths[b * (p.T + 1) * S + tid];
    for (int j = tid; j < G; j += kThreads) p.sdg[b * G + j] = sm.sdgi[j];
    __syncthreads();
  }
}

int grid_for(int64_t B) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t g = (int64_t)sms * 4;
  return (int)(B < g ? B : g);
}

}  // namespace

int launch_path_fwd_generic(const PathParams& p, cudaStream_t st) {
  size_t smem = GenericSmem::bytes(p.NL, p.H, p.n_out, p.S);
  path_fwd_generic_kernel<<<grid_for(p.B), kThreads, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int launch_path_bwd_generic(const PathParams& p, cudaStream_t st) {
  size_t smem = GenericBwdSmem::bytes(p.NL, p.H, p.n_out, p.S);
  path_bwd_generic_kernel<<<grid_for(p.B), kThreads, smem, st>>>(p);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace visde
