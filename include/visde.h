/*
 * visde.h -- C ABI of the B200-native variational path-sampling library (libvisde.so).
 *
 * The reference (Tom-Ryder/VIforSDEs) has no C ABI / FFI: its operator boundary for this
 * path is Python.  Each entry point below replaces one reference-side
 * Python operator and is what a binding for that operator would call (INTEGRATION.md shows
 * the ctypes stub a maintainer adds to the reference):
 *
 *   visde_path_fwd      <- kernels/forward.py:378-563   launch_fwd  (+ sde_fwd_kernel :91-375)
 *   visde_path_bwd      <- kernels/backward.py:627-784  launch_bwd  (+ sde_bwd_kernel :156-624)
 *   visde_elbo_fwd/_bwd <- inference/evidence_lower_bound.py:19-83 (path-dependent terms) with
 *                          inference/state_space.py:20-38, inference/types.py:19-24,
 *                          core/observations.py:52-74, examples/{ornstein_uhlenbeck,lotka_volterra}.py
 *                          (sde_kind GENERIC = user SDE: drift/diffusion tensors evaluated by the caller,
 *                          evidence_lower_bound.py:37-40, only the Gaussian algebra :77-83 runs here)
 *   visde_session_*     <- one trainer iteration's path part (inference/trainer.py:176-198) with
 *                          HOST buffers: H2D, path fwd, ELBO fwd+bwd, path bwd, D2H.
 *   visde_em_fwd/_bwd   <- core/euler_maruyama.py:11-45 (pre-training simulator, trainer.py:208-259)
 *   visde_path_summary  <- posterior/variational_posterior.py:93-135 (sample / summary)
 *   visde_grad_sqnorm, visde_adamw_ema_step <- trainer.py:199-203 (unscale, clip_grad_norm_, AdamW.step) and
 *                          exponential_moving_average.py:25-28 (EMA update)
 *
 * Conventions
 *   - plain pointers and sizes only; every device buffer is allocated by the caller; the library
 *     keeps no global mutable state (visde_last_error is thread-local); re-entrant.
 *   - all kernels are enqueued on the cudaStream_t passed as `stream` (void*; NULL = legacy default).
 *   - return 0 on success, negative VISDE_E* otherwise; visde_last_error() describes the failure.
 *   - tensors are fp32, contiguous, row-major unless a stride argument says otherwise.
 *   - weights use the nn.GRU / nn.Linear native layout (kernels/weights.py:79-196 documents the
 *     reference's transposed copies; this library reads the native tensors directly):
 *       w_ih[0] [3H, S+C+P] (input columns: state, context, theta), w_ih[k>0] [3H, H],
 *       w_hh[k] [3H, H], b_ih[k], b_hh[k] [3H]; gate row blocks r, z, n;
 *       out_w [S+S(S+1)/2, H] (rows: mu then tril row-major), out_b.
 *   - limits (kernels/constants.py:13): 1 <= NL <= 4, 1 <= H <= 256, 1 <= S <= 16.
 */
#ifndef VISDE_H
#define VISDE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VISDE_VERSION 2
#define VISDE_MAX_LAYERS 4
#define VISDE_MAX_STATE 16
#define VISDE_MAX_HIDDEN 256
#define VISDE_DIAG_MIN 1e-2f /* inference/constants.py:6 */

enum {
  VISDE_OK = 0,
  VISDE_EINVAL = -1,   /* bad dims / null pointer / unsupported config (Python raises ValueError) */
  VISDE_ECUDA = -2,    /* CUDA runtime error (Python raises RuntimeError) */
  VISDE_EWORKSPACE = -3 /* workspace too small */
};

enum { VISDE_F32 = 0, VISDE_BF16 = 1 };                 /* dtype of context / grad_context */
enum { VISDE_SDE_GENERIC = 0, VISDE_SDE_OU = 1, VISDE_SDE_LV = 2 };
/* kernel family of the recurrence: GENERIC any shape; FAST one trajectory per CTA, weights in registers;
 * TILED 4-8 trajectories per CTA, weights in shared memory (fp32 SIMT); TC 128 trajectories per CTA with the
 * gate GEMMs on tcgen05 tensor cores (fp16 hi/lo 3-pass split, FP32 accumulate).  AUTO picks by batch size. */
enum {
  VISDE_VARIANT_AUTO = 0, VISDE_VARIANT_GENERIC = 1, VISDE_VARIANT_FAST = 2, VISDE_VARIANT_TILED = 3,
  VISDE_VARIANT_TC = 4
};
/* OR-ed into visde_dims.variant: run the time-parallel GEMMs (K0/K3/K4) on the fp32 SIMT kernels
 * instead of tcgen05 3xTF32 (testing aid: the two implementations cross-check each other) */
#define VISDE_FLAG_NO_TENSOR_CORES 0x100
/* stages timed by the opt-in profiler (visde_profile_begin / _end) */
enum {
  VISDE_STAGE_K0_CTX_GEMM = 0, /* gi_ctx = ctx . W_ctx^T + b_ih_l0                */
  VISDE_STAGE_K1_PATH_FWD = 1, /* forward recurrence                               */
  VISDE_STAGE_K5_ELBO_FWD = 2,
  VISDE_STAGE_K6_ELBO_BWD = 3,
  VISDE_STAGE_K2_PATH_BWD = 4, /* reverse-time recurrence                          */
  VISDE_STAGE_K3_GRAD_CTX = 5, /* grad_context and grad_theta GEMMs                */
  VISDE_STAGE_K4_WGRAD = 6,    /* weight-gradient GEMMs + split-K reductions       */
  VISDE_NUM_STAGES = 7
};

typedef struct {
  int64_t B;  /* trajectories on this rank */
  int64_t T;  /* grid steps */
  int32_t S;  /* state dim */
  int32_t C;  /* context dim */
  int32_t P;  /* sde parameter dim */
  int32_t H;  /* GRU hidden */
  int32_t NL; /* GRU layers */
  int32_t variant; /* VISDE_VARIANT_*: kernel family selector (AUTO in production) */
} visde_dims;

typedef struct {
  const float* w_ih[VISDE_MAX_LAYERS];
  const float* w_hh[VISDE_MAX_LAYERS];
  const float* b_ih[VISDE_MAX_LAYERS];
  const float* b_hh[VISDE_MAX_LAYERS];
  const float* out_w;
  const float* out_b;
} visde_weights;

typedef struct {
  float* w_ih[VISDE_MAX_LAYERS];
  float* w_hh[VISDE_MAX_LAYERS];
  float* b_ih[VISDE_MAX_LAYERS];
  float* b_hh[VISDE_MAX_LAYERS];
  float* out_w;
  float* out_b;
} visde_weight_grads;

/* strided [B,T,C] view (the reference passes context[:, :-1] of a [B,T+1,C] tensor,
 * inference/diffusion_path_sampler.py:60-61; strides in ELEMENTS, innermost stride 1) */
typedef struct {
  const void* ptr;
  int64_t batch_stride;
  int64_t time_stride;
  int32_t dtype; /* VISDE_F32 | VISDE_BF16 */
} visde_ctx_view;

typedef struct {
  void* ptr;
  int64_t batch_stride;
  int64_t time_stride;
  int32_t dtype;
} visde_ctx_grad_view;

/* Gaussian observation model (core/observations.py:41-74) evaluated at grid indices
 * obs_idx = clamp(round(times/dt), max=T) (evidence_lower_bound.py:52; computed by the host). */
typedef struct {
  int32_t n_obs;
  int32_t obs_dim;
  const int32_t* idx;      /* device [n_obs] */
  const float* values;     /* device [n_obs, obs_dim] */
  const float* obs_matrix; /* device [obs_dim, S] or NULL (identity; requires obs_dim == S) */
  float variance;
} visde_obs;

int visde_version(void);
const char* visde_last_error(void);

/* Which kernel family visde_path_fwd (backward = 0) / visde_path_bwd (backward = 1) runs the recurrence with for these dims
 * on the current device, assuming a tcgen05-eligible context (fp32, 16-byte aligned): introspection for logs and tests, no
 * launch.  AUTO decides per direction: one trajectory per CTA below one wave of SMs, 4- / 8-trajectory tiles by a measured
 * waves x cost model, the tensor-core recurrence from B >= 3 072, the wide-state variant for S > 4.  Negative on bad dims. */
enum {
  VISDE_FAMILY_GENERIC = 0, VISDE_FAMILY_FAST = 1, VISDE_FAMILY_TILED4 = 2, VISDE_FAMILY_TILED8 = 3, VISDE_FAMILY_TC = 4,
  VISDE_FAMILY_FAST_S = 5
};
int visde_recurrence_family(const visde_dims* d, int backward);

/* bytes of the activation stash written by visde_path_fwd (save != 0) and read by _bwd */
size_t visde_stash_bytes(const visde_dims* d);
/* scratch bytes for visde_path_fwd (backward=0) / visde_path_bwd (backward=1) */
size_t visde_workspace_bytes(const visde_dims* d, int backward);

/* launch_fwd (kernels/forward.py:378): x0 [B,S] latent, theta [B,P], eps [B,T,S] ->
 * paths [B,T+1,S], means [B,T,S], chol [B,T,S,S] (upper triangle written as 0).
 * stash == NULL -> inference mode (save_activations=False, autograd.py:244-268). */
int visde_path_fwd(const visde_dims* d, float dt, const float* x0, const visde_ctx_view* ctx,
                   const float* theta, const float* eps, const visde_weights* w, float* paths,
                   float* means, float* chol, void* stash, void* workspace, size_t workspace_bytes,
                   void* stream);

/* launch_bwd (kernels/backward.py:627): cotangents g_paths [B,T+1,S], g_means [B,T,S],
 * g_chol [B,T,S,S] (lower triangle consumed, backward.py:316-324) ->
 * grad_x0 [B,S], grad_ctx (strided view, fully overwritten), grad_theta [B,P], weight grads
 * in native layout (fully overwritten; deterministic two-stage reduction, no atomics). */
int visde_path_bwd(const visde_dims* d, float dt, const float* g_paths, const float* g_means,
                   const float* g_chol, const visde_ctx_view* ctx, const float* theta,
                   const float* eps, const visde_weights* w, const float* paths, const void* stash,
                   float* grad_x0, const visde_ctx_grad_view* grad_ctx, float* grad_theta,
                   const visde_weight_grads* gw, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Path-dependent ELBO terms per trajectory: terms [B,4] = (obs, sde, gen, jacobian) log-probs
 * (evidence_lower_bound.py:28-56).  sde_kind OU/LV use built-in drift/diffusion functors
 * (examples/*.py); GENERIC reads drift [B,T,S] and diffusion [B,T,S,S] evaluated by the caller.
 * positive_mask bit s set <=> state dim s uses the softplus transform (state_space.py:20-25). */
int visde_elbo_fwd(const visde_dims* d, float dt, int sde_kind, uint32_t positive_mask,
                   const float* z, const float* means, const float* chol, const float* theta,
                   const float* drift, const float* diffusion, const visde_obs* obs, float* terms,
                   void* stream);

/* Backward of visde_elbo_fwd: g_terms [B,4] -> g_z [B,T+1,S], g_means [B,T,S], g_chol [B,T,S,S]
 * (all fully overwritten), g_theta [B,P] (OU/LV; zeroed for GENERIC), and for GENERIC
 * g_drift [B,T,S], g_diffusion [B,T,S,S] (otherwise may be NULL). */
int visde_elbo_bwd(const visde_dims* d, float dt, int sde_kind, uint32_t positive_mask,
                   const float* z, const float* means, const float* chol, const float* theta,
                   const float* drift, const float* diffusion, const visde_obs* obs,
                   const float* g_terms, float* g_z, float* g_means, float* g_chol,
                   float* g_theta, float* g_drift, float* g_diffusion, void* stream);

/* Opt-in stage profiler (bench.py's roofline): while enabled, every entry point brackets its
 * stages with CUDA events on the caller's stream; visde_profile_end synchronises those events and
 * returns the summed device time (ms) and launch count per stage.  Process-wide, off by default,
 * the only global state in the library.  max_records bounds the number of bracketed stages. */
int visde_profile_begin(int max_records);
int visde_profile_end(double* ms_per_stage, int* launches_per_stage);

/* ---- host-buffer session: one ELBO iteration of the path through HOST memory ------------- */
typedef struct visde_session visde_session;

/* Hooks of a user-defined SDE (sde_kind GENERIC): the reference calls the user's Python drift / diffusion on the
 * flattened state-space path (inference/evidence_lower_bound.py:37-40, core/sde.py:8-15) and differentiates them with
 * autograd.  The session owns the DEVICE tensors; the caller evaluates on them, on the stream it is handed, and returns
 * 0 (non-zero aborts the iteration with VISDE_EINVAL).
 *   eval: x [B*T,S] (x_t = to_state(z_t), t < T), theta [B,P] -> drift [B*T,S], diffusion [B*T,S,S] (lower-triangular)
 *   vjp : cotangents g_drift, g_diffusion -> g_x [B*T,S], g_theta [B,P] (both fully overwritten; sum over t for theta) */
typedef struct {
  int (*eval)(void* user, const float* x, const float* theta, float* drift, float* diffusion, void* stream);
  int (*vjp)(void* user, const float* x, const float* theta, const float* g_drift, const float* g_diffusion, float* g_x,
             float* g_theta, void* stream);
  void* user;
} visde_user_sde;

/* Allocates device buffers and two streams for dims d.  The host context is a dense [B,T+1,C] tensor of ctx_dtype
 * (VISDE_F32, or VISDE_BF16 = what the reference's autocast encoder emits, inference/trainer.py:171-175; halves the
 * dominant H2D term) of which rows 0..T-1 are used; grad_context comes back in the same dtype.  user_sde: hooks for
 * sde_kind GENERIC (copied), NULL otherwise. */
int visde_session_create(const visde_dims* d, int sde_kind, uint32_t positive_mask, int32_t n_obs,
                         int32_t obs_dim, int32_t ctx_dtype, const visde_user_sde* user_sde, visde_session** out);
void visde_session_destroy(visde_session* s);
/* bytes copied host->device / device->host by one visde_session_step */
size_t visde_session_h2d_bytes(const visde_session* s);
size_t visde_session_d2h_bytes(const visde_session* s);
/* number of kernels one visde_session_step launches */
int visde_session_launches(const visde_session* s);
/* eps == NULL in visde_session_step / _submit: the iteration draws its noise ON THE DEVICE, as the reference's sampler
 * does (inference/diffusion_path_sampler.py:57, torch.randn on the model's device) -- iteration i (0-based count of
 * submitted iterations) uses visde_philox_normal(seed + i, B, T, S); nothing of eps crosses the bus and
 * visde_session_h2d_bytes drops by 4 B T S.  Default seed 0. */
int visde_session_set_noise_seed(visde_session* s, uint64_t seed);

/* HOST in: x0, context [B,T+1,C] (the session's ctx_dtype), theta, eps (NULL: drawn on the device, see
 * visde_session_set_noise_seed), weights (host pointers in visde_weights),
 * obs (host pointers).  HOST out: terms [B,4], grad_x0 [B,S], grad_theta [B,P], weight grads
 * (host pointers), and grad_context [B,T+1,C] if non-NULL (row T zero).  The loss is
 * -(mean_b(obs + sde - gen + jac)); its cotangent 1/B is applied inside. Synchronous. */
int visde_session_step(visde_session* s, float dt, const float* x0, const void* context,
                       const float* theta, const float* eps, const visde_weights* w_host,
                       const visde_obs* obs_host, float* terms, float* grad_x0, float* grad_theta,
                       const visde_weight_grads* gw_host, void* grad_context);

/* Pipelined form of visde_session_step (same arguments): enqueue the H2D copies on the session's
 * copy stream and the kernels + D2H on its compute stream, and return without waiting.  The session
 * holds two device input sets, so up to TWO iterations may be in flight: the copies of iteration
 * i+1 overlap the kernels of iteration i (VISDE_EINVAL when a third is submitted).  Host input
 * buffers must stay valid, and the host output buffers of in-flight iterations distinct, until the
 * matching visde_session_wait, which blocks until the OLDEST in-flight iteration's outputs are in
 * host memory.  Pinned host memory is required for the overlap (pageable memory degrades to
 * synchronous copies).  visde_session_step == submit + wait. */
int visde_session_submit(visde_session* s, float dt, const float* x0, const void* context,
                         const float* theta, const float* eps, const visde_weights* w_host,
                         const visde_obs* obs_host, float* terms, float* grad_x0, float* grad_theta,
                         const visde_weight_grads* gw_host, void* grad_context);
int visde_session_wait(visde_session* s);

/* ---- callers either side of the path ----------------------------------------- */

/* core/euler_maruyama.py:11-45 for the built-in OU / LV functors (the theta pre-training simulator,
 * inference/trainer.py:208-259): x0 [B,S], theta [B,P] -> paths [B,T+1,S] with
 * x_{t+1} = x_t + f dt + D eps_t sqrt(dt), dims in positive_mask clamped to >= 1e-6 after every step (:41-42).
 * S, P follow sde_kind (OU 1/3, LV 2/3); VISDE_SDE_GENERIC -> VISDE_EINVAL (user SDEs are stepped in PyTorch).
 * noise [B,T,S] standard normals, or NULL: drawn in the kernel (Philox4x32-10, key = seed, counter = (t, b),
 * Box-Muller; visde_philox_normal writes the same draws), so pre-training never materialises the noise. */
int visde_em_fwd(int64_t B, int64_t T, int sde_kind, uint32_t positive_mask, float dt, const float* x0,
                 const float* theta, const float* noise, uint64_t seed, float* paths, void* stream);
/* reverse mode of visde_em_fwd (what autograd does through the reference's Python loop): g_paths [B,T+1,S] ->
 * grad_theta [B,P], grad_x0 [B,S] (may be NULL).  `paths` is the forward's output; noise / seed as in the forward. */
int visde_em_bwd(int64_t B, int64_t T, int sde_kind, uint32_t positive_mask, float dt, const float* theta,
                 const float* noise, uint64_t seed, const float* paths, const float* g_paths, float* grad_x0,
                 float* grad_theta, void* stream);
/* out [B,T,S]: counter-based standard normals -- Philox4x32-10, key = seed, counter = (t, group << 28, b), group = s / 4, Box-Muller.
 * S <= 4 is the stream visde_em_fwd / _bwd draw in-kernel; S <= 16 serves the path op as a reproducible, rank-independent replacement
 * of the reference's torch.randn(B, T, S) (inference/diffusion_path_sampler.py:57). */
int visde_philox_normal(uint64_t seed, int64_t B, int64_t T, int32_t S, float* out, void* stream);

/* posterior/variational_posterior.py:93-135 (sample / summary): z [n,T1,S] latent paths of the stash-less forward
 * -> x = from_latent(z) [n,T1,S] (may be NULL), mean [T1,S], std [T1,S] over the n samples (Bessel-corrected like
 * torch.std; nan for n == 1).  Fixed-order two-stage reduction. */
size_t visde_path_summary_workspace_bytes(int64_t n, int64_t T1, int32_t S);
int visde_path_summary(int64_t n, int64_t T1, int32_t S, uint32_t positive_mask, const float* z, float* x,
                       float* mean, float* std, void* workspace, size_t workspace_bytes, void* stream);

/* Optimiser tail of inference/trainer.py:199-203,126 over flat fp32 buffers (the all-reduce bucket):
 * visde_grad_sqnorm: sqnorm (device scalar) = [accumulate ? sqnorm : 0] + inv_scale^2 * sum g^2   (inv_scale: device
 *   scalar of GradScaler.unscale_, or NULL); deterministic.  skipped_steps (device int64 counter, may be NULL) is
 *   incremented when the result is non-finite, i.e. when some gradient is inf / NaN (GradScaler's found_inf).
 * visde_adamw_ema_step: g *= inv_scale * min(1, max_norm / (sqrt(sqnorm) + 1e-6)) (clip_grad_norm_; skipped when
 *   sqnorm == NULL or max_norm <= 0), torch.optim.AdamW update (decoupled weight decay, `step` counts from 1), then
 *   ema = lerp(ema, param, 1 - ema_decay) (exponential_moving_average.py:25-28; ema may be NULL).  When sqnorm is
 *   given and non-finite the parameter / moment update is SKIPPED like GradScaler.step() (trainer.py:202) and only the
 *   EMA moves (trainer.py:126); with skipped_steps the bias corrections use step - *skipped_steps.  No host sync. */
size_t visde_grad_sqnorm_workspace_bytes(void);
int visde_grad_sqnorm(int64_t n, const float* grads, const float* inv_scale, int accumulate, float* sqnorm,
                      int64_t* skipped_steps, void* workspace, size_t workspace_bytes, void* stream);
int visde_adamw_ema_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema,
                         float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                         float max_norm, const float* sqnorm, const float* inv_scale, float ema_decay,
                         const int64_t* skipped_steps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VISDE_H */
