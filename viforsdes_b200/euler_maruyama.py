"""``euler_maruyama`` with the signature of src/variational_sde/core/euler_maruyama.py:11-45, plus the fused
pre-training objective of inference/trainer.py:253-259.

SDEs that carry a ``device_kind`` (the built-in Ornstein-Uhlenbeck / Lotka-Volterra models of ``sde.py``) run the
whole simulation -- and, under autograd, its reverse-mode -- in ONE kernel each (``visde::em_fwd`` / ``visde::em_bwd``,
csrc/em.cu) instead of ``n_steps`` rounds of ~10 small PyTorch kernels.  ``noise=None`` draws the standard normals inside the
kernel (Philox4x32-10, reproducible from ``seed``), so the [B, T, S] noise tensor is never materialised.  User-defined
SDEs are stepped in PyTorch exactly like the reference (BASELINE.json: "generic user SDEs evaluated in PyTorch")."""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence
from typing import List, Optional

import torch
from torch import Tensor

from viforsdes_b200 import _lib
from viforsdes_b200.ops import _f32c, _ptr, _require_cuda, _stream
from viforsdes_b200.sde import SDE


def _mask(positive_dims: Sequence[int]) -> int:
    m = 0
    for d in positive_dims:
        m |= 1 << int(d)
    return m


def _kind_dims(sde_kind: int) -> tuple[int, int]:
    if sde_kind == _lib.SDE_OU:
        return 1, 3
    if sde_kind == _lib.SDE_LV:
        return 2, 3
    raise ValueError(f"sde_kind {sde_kind} has no device functor")


@torch.library.custom_op("visde::em_fwd", mutates_args=())
def em_fwd(x0: Tensor, theta: Tensor, noise: Optional[Tensor], seed: int, n_steps: int, dt: float, sde_kind: int,
           positive_mask: int) -> Tensor:
    _require_cuda(x0, theta, noise)
    S, P = _kind_dims(sde_kind)
    B = x0.shape[0]
    if tuple(x0.shape) != (B, S) or tuple(theta.shape) != (B, P):
        raise ValueError(f"x0 must be [B,{S}] and theta [B,{P}] for this SDE, got {tuple(x0.shape)}, {tuple(theta.shape)}")
    if noise is not None and tuple(noise.shape) != (B, n_steps, S):
        raise ValueError(f"noise must be [{B},{n_steps},{S}], got {tuple(noise.shape)}")
    lib = _lib.load()
    x0f, thf = _f32c(x0), _f32c(theta)
    nf = None if noise is None else _f32c(noise)
    paths = torch.empty(B, n_steps + 1, S, device=x0.device, dtype=torch.float32)
    with torch.cuda.device(x0.device):
        _lib.check(lib.visde_em_fwd(B, n_steps, sde_kind, positive_mask, dt, _ptr(x0f), _ptr(thf), _ptr(nf), seed,
                                    _ptr(paths), _stream()))
    return paths


@em_fwd.register_fake
def _(x0, theta, noise, seed, n_steps, dt, sde_kind, positive_mask):
    return torch.empty(x0.shape[0], n_steps + 1, x0.shape[1], device=x0.device, dtype=torch.float32)


@torch.library.custom_op("visde::em_bwd", mutates_args=())
def em_bwd(g_paths: Tensor, paths: Tensor, theta: Tensor, noise: Optional[Tensor], seed: int, dt: float, sde_kind: int,
           positive_mask: int) -> List[Tensor]:
    """Returns [grad_x0, grad_theta]."""
    _require_cuda(g_paths, paths, theta, noise)
    B, T1, S = paths.shape
    lib = _lib.load()
    thf, pf, gf = _f32c(theta), _f32c(paths), _f32c(g_paths)
    nf = None if noise is None else _f32c(noise)
    gx0 = torch.empty(B, S, device=paths.device, dtype=torch.float32)
    gth = torch.empty(B, theta.shape[1], device=paths.device, dtype=torch.float32)
    with torch.cuda.device(paths.device):
        _lib.check(lib.visde_em_bwd(B, T1 - 1, sde_kind, positive_mask, dt, _ptr(thf), _ptr(nf), seed, _ptr(pf), _ptr(gf),
                                    _ptr(gx0), _ptr(gth), _stream()))
    return [gx0, gth]


@em_bwd.register_fake
def _(g_paths, paths, theta, noise, seed, dt, sde_kind, positive_mask):
    f = dict(device=paths.device, dtype=torch.float32)
    return [torch.empty(paths.shape[0], paths.shape[2], **f), torch.empty(theta.shape[0], theta.shape[1], **f)]


def _em_setup(ctx, inputs, output):
    x0, theta, noise, seed, _n_steps, dt, sde_kind, positive_mask = inputs
    ctx.save_for_backward(output, theta, noise)
    ctx.consts = (seed, dt, sde_kind, positive_mask)
    ctx.dtypes = (x0.dtype, theta.dtype)


def _em_backward(ctx, g_paths):
    paths, theta, noise = ctx.saved_tensors
    seed, dt, sde_kind, positive_mask = ctx.consts
    gx0, gth = torch.ops.visde.em_bwd(g_paths.contiguous(), paths, theta, noise, seed, dt, sde_kind, positive_mask)
    return gx0.to(ctx.dtypes[0]), gth.to(ctx.dtypes[1]), None, None, None, None, None, None


em_fwd.register_autograd(_em_backward, setup_context=_em_setup)


def philox_normal(seed: int, batch: int, n_steps: int, state_dim: int, device: torch.device | str = "cuda") -> Tensor:
    """The [B, T, S] standard normals the fused simulator draws for ``seed`` (tests, reproducibility)."""
    out = torch.empty(batch, n_steps, state_dim, device=device, dtype=torch.float32)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().visde_philox_normal(seed, batch, n_steps, state_dim, _ptr(out), _stream()))
    return out


def euler_maruyama(sde: SDE, x0: Tensor, theta: Tensor, time_horizon: float, dt: float,
                   positive_dims: Sequence[int] = (), noise: Tensor | None = None, seed: int | None = None) -> Tensor:
    if dt <= 0:
        raise ValueError(f"dt must be positive, got {dt}")
    if time_horizon <= 0:
        raise ValueError(f"time_horizon must be positive, got {time_horizon}")
    n_steps = round(time_horizon / dt)
    batch, state_dim = x0.shape
    kind = getattr(sde, "device_kind", _lib.SDE_GENERIC)
    if kind != _lib.SDE_GENERIC and x0.is_cuda:
        if noise is None and seed is None:
            seed = int(torch.randint(0, 2**62, (1,)).item())  # follows torch's global generator like randn would
        out = torch.ops.visde.em_fwd(x0, theta, noise, int(seed or 0), n_steps, float(dt), kind, _mask(positive_dims))
        return out if x0.dtype == torch.float32 else out.to(x0.dtype)
    # user SDE: the reference's loop, verbatim in behaviour (core/euler_maruyama.py:27-45)
    sqrt_dt = dt**0.5
    if noise is None:
        gen = None if seed is None else torch.Generator(device=x0.device).manual_seed(seed)
        noise = torch.randn(batch, n_steps, state_dim, device=x0.device, dtype=x0.dtype, generator=gen)
    trajectory = torch.empty(batch, n_steps + 1, state_dim, device=x0.device, dtype=x0.dtype)
    trajectory[:, 0] = x0
    x = x0.clone()
    pos = list(positive_dims)
    for step in range(n_steps):
        x = x + sde.drift(x, theta) * dt + torch.einsum("bij,bj->bi", sde.diffusion(x, theta), noise[:, step]) * sqrt_dt
        if pos:
            x[:, pos] = x[:, pos].clamp(min=1e-6)
        trajectory[:, step + 1] = x
    return trajectory


def pretrain_mse(sde: SDE, theta: Tensor, obs_times: Tensor, obs_values: Tensor, time_horizon: float, dt: float,
                 positive_dims: Sequence[int] = (), noise: Tensor | None = None, seed: int | None = None) -> Tensor:
    """``Trainer._pretrain_mse_batch`` (inference/trainer.py:253-259): mean squared distance between prior simulations
    started at the first observation and the observations, at the observation grid points."""
    n = theta.shape[0]
    x0 = obs_values[0].unsqueeze(0).expand(n, -1).contiguous()
    paths = euler_maruyama(sde, x0, theta, time_horizon, dt, positive_dims, noise=noise, seed=seed)
    obs_idx = (obs_times / dt).round().long()
    return ((paths[:, obs_idx] - obs_values) ** 2).mean()


class PretrainConfig:
    """src/variational_sde/config.py:97-115 (same fields, defaults and validation)."""

    def __init__(self, n_iterations: int = 1000, batch_size: int = 4096, learning_rate: float = 0.02,
                 init_scale: float = 2.0) -> None:
        if n_iterations <= 0 or batch_size <= 0 or learning_rate <= 0 or init_scale <= 0:
            raise ValueError("value must be positive")
        self.n_iterations, self.batch_size = n_iterations, batch_size
        self.learning_rate, self.init_scale = learning_rate, init_scale


class _DeviceAdam:
    """``torch.optim.Adam`` (defaults: betas (0.9, 0.999), eps 1e-8) + ``clip_grad_norm_`` on ONE small leaf, with the whole
    update -- including "skip this step" -- expressed as tensor ops, so the pre-training loop never synchronises the host."""

    def __init__(self, param: Tensor, lr: float, max_norm: float = 1.0) -> None:
        self.p, self.lr, self.max_norm = param, lr, max_norm
        self.m, self.v = torch.zeros_like(param), torch.zeros_like(param)
        self.t = torch.zeros((), device=param.device, dtype=torch.float32)

    @torch.no_grad()
    def step(self, grad: Tensor, apply: Tensor) -> None:
        """`apply`: 0-d bool tensor; False leaves parameter and state untouched (inference/trainer.py:238-241)."""
        g = torch.nan_to_num(grad)
        coef = torch.clamp(self.max_norm / (torch.linalg.vector_norm(g) + 1e-6), max=1.0)  # nn.utils.clip_grad_norm_
        g = g * coef
        t = self.t + 1
        m = self.m.lerp(g, 0.1)
        v = self.v * 0.999 + 0.001 * g * g
        denom = v.sqrt() / torch.sqrt(1 - 0.999**t) + 1e-8
        p = self.p - (self.lr / (1 - 0.9**t)) * (m / denom)
        self.p.copy_(torch.where(apply, p, self.p))
        self.m, self.v, self.t = torch.where(apply, m, self.m), torch.where(apply, v, self.v), torch.where(apply, t, self.t)


def pretrain_sde_parameters(sde: SDE, obs_times: Tensor, obs_values: Tensor, time_horizon: float, time_step: float,
                            sde_param_positive_dims: Sequence[int], state_positive_dims: Sequence[int],
                            config: PretrainConfig | None = None, device: torch.device | str | None = None,
                            seed: int = 0) -> Tensor:
    """``Trainer.pretrain_sde_parameters`` (inference/trainer.py:208-250): fit a Gaussian over (log-)theta by Adam on the MSE
    between prior simulations and the observations; returns the mean with the best MSE seen.

    Same maths as the reference; what changes is the execution: each MSE + gradient is two fused kernels instead of
    ~25 T PyTorch kernels, the simulation noise is drawn inside the kernel (Philox, ``seed + step``), and the bookkeeping the
    reference does with three ``.item()`` calls per step (best-so-far, skip on a non-finite MSE; :235-246) is tensor algebra
    on the device, so the loop runs without a single host synchronisation."""
    cfg = config or PretrainConfig()
    dev = torch.device(device) if device is not None else obs_values.device
    d = sde.sde_param_dim
    pos = list(sde_param_positive_dims)
    init = torch.zeros(2 * d, device=dev)  # [mu, log_sigma]
    unconstrained = [i for i in range(d) if i not in pos]
    if unconstrained:
        init[unconstrained] = cfg.init_scale * torch.randn(len(unconstrained), device=dev)
    param = init.requires_grad_(True)
    opt = _DeviceAdam(param, cfg.learning_rate, max_norm=1.0)
    best_mu, best_mse = param.detach()[:d].clone(), torch.full((), float("inf"), device=dev)
    pos_mask = torch.zeros(d, dtype=torch.bool, device=dev)
    if pos:
        pos_mask[pos] = True
    obs_times, obs_values = obs_times.to(dev), obs_values.to(dev)
    for step in range(cfg.n_iterations):
        mu, log_sigma = param[:d], param[d:]
        eps = torch.randn(cfg.batch_size, d, device=dev)
        log_theta = mu + log_sigma.exp() * eps
        theta = torch.where(pos_mask, log_theta.exp(), log_theta)  # _to_constrained_theta_batch (:248-251)
        mse = pretrain_mse(sde, theta, obs_times, obs_values, time_horizon, time_step, state_positive_dims, seed=seed + step)
        finite = torch.isfinite(mse)
        better = finite & (mse.detach() < best_mse)
        best_mu = torch.where(better, mu.detach(), best_mu)
        best_mse = torch.where(better, mse.detach(), best_mse)
        (grad,) = torch.autograd.grad(mse, param)
        opt.step(grad, finite & torch.isfinite(grad).all())
    return best_mu
