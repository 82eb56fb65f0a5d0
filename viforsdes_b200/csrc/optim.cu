// Optimiser tail of one training iteration as two launches over ONE flat fp32 buffer (the gradient bucket of
// dist.FlatBucket is the same memory the all-reduce just averaged): global gradient norm, then clip + AdamW + EMA.
//
// Restates inference/trainer.py:199-203 + :126 for fp32 parameters:
//   scaler.unscale_ (inv_scale), nn.utils.clip_grad_norm_(params, max_norm)   coef = min(1, max_norm / (norm + 1e-6))
//   torch.optim.AdamW.step (decoupled weight decay, bias-corrected, eps outside the sqrt of the corrected v)
//   ExponentialMovingAverage.update: shadow.lerp_(param, 1 - decay)          (exponential_moving_average.py:25-28)
// The reference runs these as ~3 foreach passes per tensor list plus two .item() syncs; here the
// clip coefficient stays on the device and every element is read and written once: 5 streams read, 4 written.
// Reductions are two-stage in a fixed order (per-CTA partial, then one warp sums the partials): deterministic.
#include "common.cuh"

namespace visde {
namespace {

constexpr int kOptThreads = 256, kOptMaxBlocks = 1184;  // 8 CTAs per SM x 148

__global__ void __launch_bounds__(kOptThreads) sqnorm_partial_kernel(const float* __restrict__ g, int64_t n, float* part) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * kOptThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kOptThreads + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    } else {
      for (int64_t j = i; j < n; ++j) acc = fmaf(g[j], g[j], acc);
    }
  }
  __shared__ float red[kOptThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kOptThreads / 32; ++i) s += red[i];
    part[blockIdx.x] = s;
  }
}

// one warp: sums the partials in a fixed order; sqnorm = (accumulate ? sqnorm : 0) + sum * inv_scale^2
__global__ void sqnorm_final_kernel(const float* part, int nparts, const float* inv_scale, int accumulate, float* sqnorm,
                                    long long* skipped) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 32) acc += part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) {
    const float is = inv_scale ? *inv_scale : 1.f;
    const float r = (accumulate ? *sqnorm : 0.f) + acc * is * is;
    *sqnorm = r;
    // GradScaler.step (trainer.py:202): a non-finite gradient anywhere makes the squared norm non-finite; the optimiser
    // step is then skipped and does not count towards the bias corrections
    if (skipped && !isfinite(r)) *skipped += 1;
  }
}

struct AdamParams {
  int64_t n;
  float* p;
  const float* g;
  float* m;
  float* v;
  float* ema;  // or nullptr
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, max_norm, ema_w;
  const float* sqnorm;     // device scalar: squared global gradient norm (after unscale), or nullptr = no clipping
  const float* inv_scale;  // device scalar (GradScaler), or nullptr
  const long long* skipped;  // device counter of skipped (non-finite) steps, or nullptr
  long long step;
};

__device__ __forceinline__ void adam_elem(const AdamParams& a, float coef, float& p, float g, float& m, float& v, float& e) {
  g *= coef;
  p = p * (1.f - a.lr * a.wd);
  m = fmaf(1.f - a.beta1, g - m, m);            // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.beta2, v, (1.f - a.beta2) * g * g);  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
  e = fmaf(a.ema_w, p - e, e);                  // shadow.lerp_(param, 1 - decay)
}

// the skipped step of GradScaler.step(): parameters and moments stay, the EMA shadow still moves (trainer.py:126)
__global__ void __launch_bounds__(kOptThreads) adamw_ema_kernel(AdamParams a) {
  if (a.sqnorm && !isfinite(*a.sqnorm)) {  // found_inf: scaler.step() skips optimizer.step()
    if (!a.ema) return;
    const int64_t stride = (int64_t)gridDim.x * kOptThreads;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < a.n; i += stride)
      a.ema[i] = fmaf(a.ema_w, a.p[i] - a.ema[i], a.ema[i]);
    return;
  }
  if (a.skipped) {  // bias corrections from the number of steps actually taken
    const double s = (double)(a.step - *a.skipped);
    a.bc1 = (float)(1.0 - pow((double)a.beta1, s));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, s));
  }
  float coef = a.inv_scale ? *a.inv_scale : 1.f;
  if (a.sqnorm && a.max_norm > 0.f) {
    const float c = a.max_norm / (sqrtf(*a.sqnorm) + 1e-6f);  // nn.utils.clip_grad_norm_
    coef *= fminf(c, 1.f);
  }
  const int64_t stride = (int64_t)gridDim.x * kOptThreads * 4;
  for (int64_t i = ((int64_t)blockIdx.x * kOptThreads + threadIdx.x) * 4; i < a.n; i += stride) {
    if (i + 3 < a.n) {
      float4 p = *reinterpret_cast<float4*>(a.p + i), m = *reinterpret_cast<float4*>(a.m + i), v = *reinterpret_cast<float4*>(a.v + i);
      const float4 g = *reinterpret_cast<const float4*>(a.g + i);
      float4 e = a.ema ? *reinterpret_cast<float4*>(a.ema + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      adam_elem(a, coef, p.x, g.x, m.x, v.x, e.x);
      adam_elem(a, coef, p.y, g.y, m.y, v.y, e.y);
      adam_elem(a, coef, p.z, g.z, m.z, v.z, e.z);
      adam_elem(a, coef, p.w, g.w, m.w, v.w, e.w);
      *reinterpret_cast<float4*>(a.p + i) = p;
      *reinterpret_cast<float4*>(a.m + i) = m;
      *reinterpret_cast<float4*>(a.v + i) = v;
      if (a.ema) *reinterpret_cast<float4*>(a.ema + i) = e;
    } else {
      for (int64_t j = i; j < a.n; ++j) {
        float p = a.p[j], m = a.m[j], v = a.v[j], e = a.ema ? a.ema[j] : 0.f;
        adam_elem(a, coef, p, a.g[j], m, v, e);
        a.p[j] = p; a.m[j] = m; a.v[j] = v;
        if (a.ema) a.ema[j] = e;
      }
    }
  }
}

int opt_blocks(int64_t n) {
  int64_t b = (n + kOptThreads * 4 - 1) / (kOptThreads * 4);
  return (int)(b < 1 ? 1 : b > kOptMaxBlocks ? kOptMaxBlocks : b);
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace visde

using namespace visde;

extern "C" {

size_t visde_grad_sqnorm_workspace_bytes(void) { return kOptMaxBlocks * sizeof(float); }

int visde_grad_sqnorm(int64_t n, const float* grads, const float* inv_scale, int accumulate, float* sqnorm,
                      int64_t* skipped_steps, void* workspace, size_t workspace_bytes, void* stream) {
  VISDE_REQUIRE(n >= 0, "grad_sqnorm: negative size");
  VISDE_REQUIRE(sqnorm, "grad_sqnorm: sqnorm is NULL");
  VISDE_REQUIRE(n == 0 || (grads && aligned16(grads)), "grad_sqnorm: grads must be a 16-byte aligned device pointer");
  if (workspace_bytes < visde_grad_sqnorm_workspace_bytes() || !workspace) {
    set_error("grad_sqnorm: workspace too small (%zu < %zu)", workspace_bytes, visde_grad_sqnorm_workspace_bytes());
    return VISDE_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = opt_blocks(n);
  sqnorm_partial_kernel<<<nb, kOptThreads, 0, st>>>(grads, n, reinterpret_cast<float*>(workspace));
  VISDE_CUDA_CHECK(cudaGetLastError());
  sqnorm_final_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const float*>(workspace), nb, inv_scale, accumulate, sqnorm,
                                        reinterpret_cast<long long*>(skipped_steps));
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

int visde_adamw_ema_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema,
                         float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                         float max_norm, const float* sqnorm, const float* inv_scale, float ema_decay,
                         const int64_t* skipped_steps, void* stream) {
  VISDE_REQUIRE(n >= 0, "adamw_ema_step: negative size");
  VISDE_REQUIRE(step >= 1, "adamw_ema_step: step counts from 1, got %lld", (long long)step);
  VISDE_REQUIRE(lr >= 0.f && eps >= 0.f && weight_decay >= 0.f, "adamw_ema_step: invalid learning rate / eps / weight decay");
  VISDE_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "adamw_ema_step: betas must be in [0, 1)");
  if (n == 0) return VISDE_OK;
  VISDE_REQUIRE(params && grads && exp_avg && exp_avg_sq, "adamw_ema_step: NULL tensor argument");
  VISDE_REQUIRE(aligned16(params) && aligned16(grads) && aligned16(exp_avg) && aligned16(exp_avg_sq) && aligned16(ema),
                "adamw_ema_step: buffers must be 16-byte aligned");
  AdamParams a{};
  a.n = n; a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.ema = ema;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.max_norm = max_norm; a.ema_w = 1.f - ema_decay; a.sqnorm = sqnorm; a.inv_scale = inv_scale;
  a.skipped = reinterpret_cast<const long long*>(skipped_steps); a.step = step;
  adamw_ema_kernel<<<opt_blocks(n), kOptThreads, 0, (cudaStream_t)stream>>>(a);
  VISDE_CUDA_CHECK(cudaGetLastError());
  return VISDE_OK;
}

}  // extern "C"
