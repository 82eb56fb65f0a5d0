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
  __host__ __device__ static size_t bytes(int NL, int H, int n_out, int S) {
    return sizeof(float) * (size_t)(NL * H + 9 * H + n_out + 2 * S);
  }
};

// SW: the recurrent, output and state-column weights are staged once per CTA in shared memory, transposed
// ([k][row], padded row pitch) so that thread `row` walks its dot product with conflict-free LDS instead of a
// strided L2 stream (18 ms -> ~1 ms per launch at B = 1024, T = 100, S = 10).  Used whenever they fit (H = 64, NL = 2,
// S = 10: 171 KB); larger shapes stream from L2 as before.
template <bool SW, int HT>
__global__ void __launch_bounds__(kThreads) path_fwd_generic_kernel(PathParams p) {
  extern __shared__ float smem_f[];
  const int H = HT > 0 ? HT : p.H, S = p.S, NL = p.NL, G = 3 * H;  // HT > 0: compile-time trip counts
  const int ld0 = p.S + p.C + p.P;
  GenericSmem sm(smem_f, NL, H, p.n_out, S);
  const int tid = threadIdx.x;
  const int GP = G + 1, OP = p.n_out + 1;  // padded pitches of the transposed copies
  float* wt_hh = smem_f + GenericSmem::bytes(NL, H, p.n_out, S) / sizeof(float);  // [NL][H][GP]
  float* wt_ih = wt_hh + (size_t)NL * H * GP;                                      // [NL-1][H][GP]
  float* wt_out = wt_ih + (size_t)(NL - 1) * H * GP;                               // [H][OP]
  float* wt_z = wt_out + (size_t)H * OP;                                           // [S][GP]
  if (SW) {
    for (int k = 0; k < NL; ++k)
      for (int idx = tid; idx < G * H; idx += kThreads) {
        const int j = idx / H, q = idx % H;
        wt_hh[((size_t)k * H + q) * GP + j] = p.w_hh[k][idx];
        if (k > 0) wt_ih[((size_t)(k - 1) * H + q) * GP + j] = p.w_ih[k][idx];
      }
    for (int idx = tid; idx < p.n_out * H; idx += kThreads) wt_out[(size_t)(idx % H) * OP + idx / H] = p.out_w[idx];
    for (int idx = tid; idx < G * S; idx += kThreads) wt_z[(size_t)(idx % S) * GP + idx / S] = p.w_ih[0][(int64_t)(idx / S) * ld0 + idx % S];
    __syncthreads();
  }

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

    // one gate row per thread when 3H <= kThreads: the layer-0 input of the next step is prefetched a step ahead
    const bool pre = G <= kThreads;
    float gi_pre = (pre && tid < G && p.T > 0) ? p.gi_ctx[b * p.T * G + tid] : 0.f;
    for (int64_t t = 0; t < p.T; ++t) {
      const int64_t row = b * p.T + t;
      const float gi_now = gi_pre;
      if (pre && tid < G && t + 1 < p.T) gi_pre = p.gi_ctx[(row + 1) * G + tid];
      for (int k = 0; k < NL; ++k) {
        // gate pre-activations, one row per thread
        for (int j = tid; j < G; j += kThreads) {
          float gi, gh = p.b_hh[k][j];
          const float* hk = sm.h + k * H;
          if (SW) {
            const float* w = wt_hh + (size_t)k * H * GP + j;
            float g2 = 0.f;
#pragma unroll 8
            for (int q = 0; q + 1 < H; q += 2) {
              gh = fmaf(w[q * GP], hk[q], gh);
              g2 = fmaf(w[(q + 1) * GP], hk[q + 1], g2);
            }
            if (H & 1) g2 = fmaf(w[(H - 1) * GP], hk[H - 1], g2);
            gh += g2;
          } else {
            const float* whh = p.w_hh[k] + (int64_t)j * H;
            for (int q = 0; q < H; ++q) gh += whh[q] * hk[q];
          }
          if (k == 0) {
            gi = (pre ? gi_now : p.gi_ctx[row * G + j]) + sm.gth[j];
            if (SW) {
              for (int s = 0; s < S; ++s) gi = fmaf(wt_z[(size_t)s * GP + j], sm.z[s], gi);
            } else {
              const float* wz = p.w_ih[0] + (int64_t)j * ld0;
              for (int s = 0; s < S; ++s) gi += wz[s] * sm.z[s];
            }
          } else {
            gi = p.b_ih[k][j];
            const float* hb = sm.h + (k - 1) * H;
            if (SW) {
              const float* w = wt_ih + (size_t)(k - 1) * H * GP + j;
              float g2 = 0.f;
#pragma unroll 8
              for (int q = 0; q + 1 < H; q += 2) {
                gi = fmaf(w[q * GP], hb[q], gi);
                g2 = fmaf(w[(q + 1) * GP], hb[q + 1], g2);
              }
              if (H & 1) g2 = fmaf(w[(H - 1) * GP], hb[H - 1], g2);
              gi += g2;
            } else {
              const float* wih = p.w_ih[k] + (int64_t)j * H;
              for (int q = 0; q < H; ++q) gi += wih[q] * hb[q];
            }
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
        const float* ht = sm.h + (NL - 1) * H;
        if (SW) {
          float a2 = 0.f;
#pragma unroll 8
          for (int q = 0; q + 1 < H; q += 2) {
            acc = fmaf(wt_out[q * OP + m], ht[q], acc);
            a2 = fmaf(wt_out[(q + 1) * OP + m], ht[q + 1], a2);
          }
          if (H & 1) a2 = fmaf(wt_out[(H - 1) * OP + m], ht[H - 1], a2);
          acc += a2;
        } else {
          const float* wo = p.out_w + (int64_t)m * H;
          for (int q = 0; q < H; ++q) acc += wo[q] * ht[q];
        }
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
  __host__ __device__ static size_t bytes(int NL, int H, int n_out, int S) {
    return sizeof(float) * (size_t)(NL * H + 8 * H + n_out + 2 * S);
  }
};

// SW: weights staged once per CTA in shared memory (native row-major: thread i reads column i, conflict-free).
// In both modes the transposed products are split over kThreads / H row ranges and summed through shared memory.
template <bool SW, int HT>
__global__ void __launch_bounds__(kThreads) path_bwd_generic_kernel(PathParams p) {
  extern __shared__ float smem_f[];
  const int H = HT > 0 ? HT : p.H, S = p.S, NL = p.NL, G = 3 * H;
  const int ld0 = p.S + p.C + p.P;
  GenericBwdSmem sm(smem_f, NL, H, p.n_out, S);
  const int tid = threadIdx.x;
  const int64_t srow = stash_row_floats(NL, H);
  const int nparts = kThreads / H > 0 ? kThreads / H : 1;  // row ranges of the transposed products (H <= 256)
  float* red = smem_f + GenericBwdSmem::bytes(NL, H, p.n_out, S) / sizeof(float);  // [2][nparts][H] + [16 parts][S]
  float* redz = red + 2 * nparts * H;
  float* s_whh = redz + 16 * S;                        // [NL][G][H]
  float* s_wih = s_whh + (size_t)NL * G * H;           // [NL-1][G][H]
  float* s_wout = s_wih + (size_t)(NL - 1) * G * H;    // [n_out][H]
  float* s_wz = s_wout + (size_t)p.n_out * H;          // [G][S]
  if (SW) {
    for (int k = 0; k < NL; ++k)
      for (int idx = tid; idx < G * H; idx += kThreads) {
        s_whh[(size_t)k * G * H + idx] = p.w_hh[k][idx];
        if (k > 0) s_wih[(size_t)(k - 1) * G * H + idx] = p.w_ih[k][idx];
      }
    for (int idx = tid; idx < p.n_out * H; idx += kThreads) s_wout[idx] = p.out_w[idx];
    for (int idx = tid; idx < G * S; idx += kThreads) s_wz[idx] = p.w_ih[0][(int64_t)(idx / S) * ld0 + idx % S];
    __syncthreads();
  }

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
      if (tid < nparts * H) {
        // W_out^T d_out, split over row ranges like the recurrent products
        const int i = tid % H, part = tid / H;
        const int m0 = part * p.n_out / nparts, m1 = (part + 1) * p.n_out / nparts;
        const float* wo = SW ? s_wout : p.out_w;
        float acc = 0.f;
        for (int m = m0; m < m1; ++m) acc += wo[(int64_t)m * H + i] * sm.dout[m];
        red[part * H + i] = acc;
      }
      __syncthreads();
      for (int i = tid; i < H; i += kThreads) {
        float acc = sm.dhc[(NL - 1) * H + i];
        for (int q = 0; q < nparts; ++q) acc += red[q * H + i];
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
        if (tid < nparts * H) {
          // partial transposed products over this thread's range of gate rows
          const int i = tid % H, part = tid / H;
          const int j0 = part * H / nparts, j1 = (part + 1) * H / nparts;
          const float* whh = SW ? s_whh + (size_t)k * G * H : p.w_hh[k];
          float acc = 0.f, accb = 0.f;
#pragma unroll 4
          for (int j = j0; j < j1; ++j) {
            acc += whh[(int64_t)(j)*H + i] * sm.dg[j];
            acc += whh[(int64_t)(H + j) * H + i] * sm.dg[H + j];
            acc += whh[(int64_t)(2 * H + j) * H + i] * sm.dg[3 * H + j];
          }
          if (k > 0) {
            const float* wih = SW ? s_wih + (size_t)(k - 1) * G * H : p.w_ih[k];
#pragma unroll 4
            for (int j = j0; j < j1; ++j) {
              accb += wih[(int64_t)j * H + i] * sm.dg[j];
              accb += wih[(int64_t)(H + j) * H + i] * sm.dg[H + j];
              accb += wih[(int64_t)(2 * H + j) * H + i] * sm.dg[2 * H + j];
            }
          }
          red[part * H + i] = acc;
          red[(nparts + part) * H + i] = accb;
        }
        if (k == 0) {
          // d z_t += W_ih_l0[:, :S]^T d_gi: 16 row ranges x S columns
          const int s = tid % 16, part = tid / 16;
          if (s < S) {
            const int j0 = part * G / 16, j1 = (part + 1) * G / 16;
            float acc = 0.f;
            if (SW) {
              for (int j = j0; j < j1; ++j) acc += s_wz[j * S + s] * sm.dg[j];
            } else {
              for (int j = j0; j < j1; ++j) acc += p.w_ih[0][(int64_t)j * ld0 + s] * sm.dg[j];
            }
            redz[part * S + s] = acc;
          }
          for (int j = tid; j < G; j += kThreads) sm.sdgi[j] += sm.dg[j];
        }
        __syncthreads();
        for (int i = tid; i < H; i += kThreads) {
          float acc = 0.f, accb = 0.f;
          for (int q = 0; q < nparts; ++q) {
            acc += red[q * H + i];
            accb += red[(nparts + q) * H + i];
          }
          sm.dhc[k * H + i] += acc;
          if (k > 0) sm.dh[i] = sm.dhc[(k - 1) * H + i] + accb;
        }
        if (k == 0 && tid < S) {
          float acc = 0.f;
          for (int q = 0; q < 16; ++q) acc += redz[q * S + tid];
          sm.dz[tid] += acc;
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
  const size_t base = GenericSmem::bytes(p.NL, p.H, p.n_out, p.S);
  const int G = 3 * p.H;
  const size_t wfl = (size_t)p.NL * p.H * (G + 1) + (size_t)(p.NL - 1) * p.H * (G + 1) + (size_t)p.H * (p.n_out + 1) +
                     (size_t)p.S * (G + 1);
  const size_t smem_w = base + sizeof(float) * wfl;
  if (smem_w <= 227 * 1024) {
    static DeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.needed(&attr_dev)) {
      VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_fwd_generic_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_fwd_generic_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_once.done(attr_dev);
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(p.B < sms ? p.B : sms);
    if (p.H == 64)
      path_fwd_generic_kernel<true, 64><<<grid, kThreads, smem_w, st>>>(p);
    else
      path_fwd_generic_kernel<true, 0><<<grid, kThreads, smem_w, st>>>(p);
  } else {
    path_fwd_generic_kernel<false, 0><<<grid_for(p.B), kThreads, base, st>>>(p);
  }
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int launch_path_bwd_generic(const PathParams& p, cudaStream_t st) {
  const int G = 3 * p.H;
  const int nparts = kThreads / p.H > 0 ? kThreads / p.H : 1;
  const size_t base = GenericBwdSmem::bytes(p.NL, p.H, p.n_out, p.S) + sizeof(float) * ((size_t)2 * nparts * p.H + 16 * p.S);
  const size_t wfl = (size_t)p.NL * G * p.H + (size_t)(p.NL - 1) * G * p.H + (size_t)p.n_out * p.H + (size_t)G * p.S;
  const size_t smem_w = base + sizeof(float) * wfl;
  if (smem_w <= 227 * 1024) {
    static DeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.needed(&attr_dev)) {
      VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_generic_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      VISDE_CUDA_CHECK(cudaFuncSetAttribute(path_bwd_generic_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_once.done(attr_dev);
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(p.B < sms ? p.B : sms);
    if (p.H == 64)
      path_bwd_generic_kernel<true, 64><<<grid, kThreads, smem_w, st>>>(p);
    else
      path_bwd_generic_kernel<true, 0><<<grid, kThreads, smem_w, st>>>(p);
  } else {
    path_bwd_generic_kernel<false, 0><<<grid_for(p.B), kThreads, base, st>>>(p);
  }
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // namespace visde
