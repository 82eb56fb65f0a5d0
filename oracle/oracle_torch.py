"""CPU oracle for the variational path-sampling hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``viforsdes_b200/``
imports it, and the product path raises when the CUDA library is missing.

It restates, in plain differentiable PyTorch (CPU, fp32 or fp64), the algorithm of
Tom-Ryder/VIforSDEs for the path ``BASELINE.json:north_star`` names.  Every function
cites the reference ``file:line`` (relative to ``/root/reference``) it follows.

Parity pinning: the reference ships no golden vectors or tests, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (imports ``/root/reference/src``; runs
``DiffusionTransitionHead.forward`` stepwise, ``compute_evidence_lower_bound`` and the
reference Triton kernels under ``TRITON_INTERPRET=1``) and committed as
``tests/golden/*.pt``.  ``tests/test_oracle_golden.py`` checks this file against them.

Third-party arithmetic at the boundary (torch, pinned 2.9.1 in the reference's
``uv.lock:1225``; 2.11.0 here): ``nn.GRU`` cell semantics (restated by hand below),
``F.softplus`` (threshold 20), ``F.logsigmoid``, ``MultivariateNormal.log_prob``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Sequence

import torch
from torch import Tensor
from torch.distributions import MultivariateNormal
from torch.nn import functional as F

DIAG_MIN = 1e-2  # src/variational_sde/inference/constants.py:6
MAX_LAYERS = 4  # src/variational_sde/kernels/constants.py:13


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
@dataclass
class HeadWeights:
    """Head parameters in ``nn.GRU`` / ``nn.Linear`` native layout.

    ``w_ih[0]``: [3H, S+C+P] with input columns ordered (state, context, theta)
    (``models/head.py:75`` concatenation order); ``w_ih[k>0]``: [3H, H]; ``w_hh[k]``:
    [3H, H]; biases [3H]; gate row blocks ordered r, z, n (``kernels/constants.py:7-10``);
    ``out_w``: [S + n_tril, H] rows = mu then tril row-major (``models/head.py:80-97``).
    """

    w_ih: list[Tensor]
    w_hh: list[Tensor]
    b_ih: list[Tensor]
    b_hh: list[Tensor]
    out_w: Tensor
    out_b: Tensor
    state_dim: int
    context_dim: int
    param_dim: int

    @property
    def hidden_dim(self) -> int:
        return self.w_hh[0].shape[1]

    @property
    def num_layers(self) -> int:
        return len(self.w_hh)

    def tensors(self) -> list[Tensor]:
        return [*self.w_ih, *self.w_hh, *self.b_ih, *self.b_hh, self.out_w, self.out_b]

    def map(self, fn: Callable[[Tensor], Tensor]) -> "HeadWeights":
        return HeadWeights(
            [fn(t) for t in self.w_ih],
            [fn(t) for t in self.w_hh],
            [fn(t) for t in self.b_ih],
            [fn(t) for t in self.b_hh],
            fn(self.out_w),
            fn(self.out_b),
            self.state_dim,
            self.context_dim,
            self.param_dim,
        )


def make_head_weights(
    state_dim: int,
    context_dim: int,
    param_dim: int,
    hidden_dim: int,
    num_layers: int,
    *,
    seed: int = 0,
    out_scale: float = 0.1,
    dtype: torch.dtype = torch.float32,
) -> HeadWeights:
    """Default ``nn.GRU`` init U(+-1/sqrt(H)); ``out_proj`` bias as ``models/head.py:60-66``
    (zeros, 1.0 on the Cholesky diagonal) with weight ~ N(0, out_scale) instead of the
    reference's zero init (zero init makes every GRU gradient vanish)."""
    if not 1 <= num_layers <= MAX_LAYERS:
        raise ValueError(f"num_layers must be in [1, {MAX_LAYERS}], got {num_layers}")
    g = torch.Generator().manual_seed(seed)
    H, S = hidden_dim, state_dim
    k = 1.0 / math.sqrt(H)

    def u(*shape: int) -> Tensor:
        return ((torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1) * k).to(dtype)

    w_ih = [u(3 * H, S + context_dim + param_dim)] + [u(3 * H, H) for _ in range(num_layers - 1)]
    w_hh = [u(3 * H, H) for _ in range(num_layers)]
    b_ih = [u(3 * H) for _ in range(num_layers)]
    b_hh = [u(3 * H) for _ in range(num_layers)]
    n_tril = S * (S + 1) // 2
    out_w = (torch.randn(S + n_tril, H, generator=g, dtype=torch.float64) * out_scale).to(dtype)
    out_b = torch.zeros(S + n_tril, dtype=dtype)
    for d in range(S):
        out_b[S + d * (d + 3) // 2] = 1.0  # models/head.py:64-66
    return HeadWeights(w_ih, w_hh, b_ih, b_hh, out_w, out_b, S, context_dim, param_dim)


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
class _LowerBound(torch.autograd.Function):
    """``max(x, bound)`` whose gradient passes iff ``x >= bound`` or ``grad < 0``
    (src/variational_sde/primitives/bounds.py:10-24)."""

    @staticmethod
    def forward(ctx, x: Tensor, bound: Tensor) -> Tensor:  # type: ignore[override]
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output: Tensor):  # type: ignore[override]
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (grad_output < 0)) * grad_output, None


def lower_bound(x: Tensor, bound: float) -> Tensor:
    return _LowerBound.apply(x, torch.as_tensor(bound, dtype=x.dtype, device=x.device))


def tril_from_params(params: Tensor, state_dim: int) -> Tensor:
    """Row-major lower-triangular fill with the diagonal floored at DIAG_MIN
    (src/variational_sde/models/head.py:88-97)."""
    rows, cols = torch.tril_indices(state_dim, state_dim)
    diag = rows == cols
    L = torch.zeros(params.shape[0], state_dim, state_dim, dtype=params.dtype)
    L[:, rows[~diag], cols[~diag]] = params[:, ~diag]
    L[:, rows[diag], cols[diag]] = lower_bound(params[:, diag], DIAG_MIN)
    return L


def gru_cell(x: Tensor, h: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """One ``nn.GRU`` layer step (torch semantics used by models/head.py:78 and restated
    by the reference kernel at kernels/forward.py:33-88): gates r, z, n;
    ``n = tanh(W_in x + b_in + r * (W_hn h + b_hn))``; ``h' = (1 - z) n + z h``."""
    H = h.shape[-1]
    gi = x @ w_ih.T + b_ih
    gh = h @ w_hh.T + b_hh
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H : 2 * H] + gh[:, H : 2 * H])
    n = torch.tanh(gi[:, 2 * H :] + r * gh[:, 2 * H :])
    return (1.0 - z) * n + z * h


def head_step(w: HeadWeights, z_t: Tensor, ctx_t: Tensor, theta: Tensor, hidden: list[Tensor]):
    """``DiffusionTransitionHead.forward`` (src/variational_sde/models/head.py:68-86)."""
    x = torch.cat([z_t, ctx_t, theta], dim=-1)
    new_hidden = []
    for k in range(w.num_layers):
        hk = gru_cell(x, hidden[k], w.w_ih[k], w.w_hh[k], w.b_ih[k], w.b_hh[k])
        new_hidden.append(hk)
        x = hk
    params = x @ w.out_w.T + w.out_b
    S = w.state_dim
    mu = params[:, :S]
    L = tril_from_params(params[:, S:], S)
    return mu, L, new_hidden


def sample_paths(w: HeadWeights, x0: Tensor, context: Tensor, theta: Tensor, eps: Tensor, dt: float):
    """Sequential reparameterised path sampling: the loop the reference kernel fuses
    (src/variational_sde/kernels/forward.py:192-375; step spec models/head.py:68-86).

    x0 [B,S] (latent z0), context [B,T,C], theta [B,P], eps [B,T,S] ->
    paths [B,T+1,S], means [B,T,S], chol [B,T,S,S] (upper triangle zero)."""
    B, T = context.shape[0], context.shape[1]
    dtype = x0.dtype
    hidden = [torch.zeros(B, w.hidden_dim, dtype=dtype) for _ in range(w.num_layers)]
    z = x0
    paths, means, chols = [z], [], []
    sqrt_dt = math.sqrt(dt)
    for t in range(T):
        mu, L, hidden = head_step(w, z, context[:, t], theta, hidden)
        z = z + mu * dt + torch.einsum("bij,bj->bi", L, eps[:, t]) * sqrt_dt
        paths.append(z)
        means.append(mu)
        chols.append(L)
    return torch.stack(paths, 1), torch.stack(means, 1), torch.stack(chols, 1)


# --------------------------------------------------------------------------------------
# state space (src/variational_sde/inference/state_space.py:20-38, types.py:19-24)
# --------------------------------------------------------------------------------------
def to_state(z: Tensor, positive_dims: Sequence[int]) -> Tensor:
    if not positive_dims:
        return z
    x = z.clone()
    x[..., list(positive_dims)] = F.softplus(z[..., list(positive_dims)])
    return x


def to_latent(x: Tensor, positive_dims: Sequence[int]) -> Tensor:
    if not positive_dims:
        return x
    z = x.clone()
    xp = x[..., list(positive_dims)].clamp(min=1e-6)
    z[..., list(positive_dims)] = xp + torch.log(-torch.expm1(-xp))
    return z


def log_jacobian(z: Tensor, positive_dims: Sequence[int]) -> Tensor:
    """sum_{t>=1} sum_{pos} logsigmoid(z_t) (types.py:23-24 -> state_space.py:35-38)."""
    if not positive_dims:
        return torch.zeros(z.shape[0], dtype=z.dtype)
    return F.logsigmoid(z[:, 1:][..., list(positive_dims)]).sum(dim=-1).sum(dim=-1)


# --------------------------------------------------------------------------------------
# SDEs (protocol: src/variational_sde/core/sde.py:8-15)
# --------------------------------------------------------------------------------------
class OrnsteinUhlenbeck:
    """examples/ornstein_uhlenbeck.py:18-30."""

    state_dim = 1
    sde_param_dim = 3
    positive_dims: tuple[int, ...] = ()

    def drift(self, x: Tensor, p: Tensor) -> Tensor:
        return p[..., 0:1] * (p[..., 1:2] - x)

    def diffusion(self, x: Tensor, p: Tensor) -> Tensor:
        return p[..., 2:3].reshape(x.shape[0], 1, 1)


class LotkaVolterra:
    """examples/lotka_volterra.py:18-46 (Cholesky of the 2x2 diffusion matrix, 3 clamps)."""

    state_dim = 2
    sde_param_dim = 3
    positive_dims: tuple[int, ...] = (0, 1)

    def drift(self, x: Tensor, p: Tensor) -> Tensor:
        u, v = x[..., 0], x[..., 1]
        t1, t2, t3 = p[..., 0], p[..., 1], p[..., 2]
        return torch.stack([t1 * u - t2 * u * v, t2 * u * v - t3 * v], dim=-1)

    def diffusion(self, x: Tensor, p: Tensor) -> Tensor:
        u, v = x[..., 0], x[..., 1]
        t1, t2, t3 = p[..., 0], p[..., 1], p[..., 2]
        uv = u * v
        b11 = t1 * u + t2 * uv
        b12 = -t2 * uv
        b22 = t3 * v + t2 * uv
        L00 = torch.sqrt(b11.clamp(min=1e-6))
        L10 = b12 / L00.clamp(min=1e-6)
        L11 = torch.sqrt((b22 - L10**2).clamp(min=1e-6))
        zeros = torch.zeros_like(L00)
        return torch.stack([torch.stack([L00, zeros], -1), torch.stack([L10, L11], -1)], -2)


class Lorenz96:
    """User-defined SDE of BASELINE.json config 5: cyclic
    ``drift_i = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + F``, ``diffusion = sigma I``."""

    sde_param_dim = 2
    positive_dims: tuple[int, ...] = ()

    def __init__(self, state_dim: int = 10) -> None:
        self.state_dim = state_dim

    def drift(self, x: Tensor, p: Tensor) -> Tensor:
        xp1 = torch.roll(x, -1, dims=-1)
        xm1 = torch.roll(x, 1, dims=-1)
        xm2 = torch.roll(x, 2, dims=-1)
        return (xp1 - xm2) * xm1 - x + p[..., 0:1]

    def diffusion(self, x: Tensor, p: Tensor) -> Tensor:
        eye = torch.eye(self.state_dim, dtype=x.dtype, device=x.device)
        return p[..., 1].reshape(-1, 1, 1) * eye


# --------------------------------------------------------------------------------------
# ELBO terms (src/variational_sde/inference/evidence_lower_bound.py:19-83)
# --------------------------------------------------------------------------------------
def gaussian_log_prob(x: Tensor, mu: Tensor, L: Tensor) -> Tensor:
    """evidence_lower_bound.py:77-83: MVN(scale_tril) over (B*T) rows, summed over t."""
    B = x.shape[0]
    S = x.shape[-1]
    dist = MultivariateNormal(loc=mu.reshape(-1, S), scale_tril=L.reshape(-1, S, S))
    return dist.log_prob(x.reshape(-1, S)).reshape(B, -1).sum(dim=-1)


def gaussian_obs_log_prob(obs: Tensor, state: Tensor, variance: float, obs_matrix: Tensor | None) -> Tensor:
    """src/variational_sde/core/observations.py:52-74."""
    pred = torch.einsum("od,...d->...o", obs_matrix, state) if obs_matrix is not None else state
    diff = obs - pred
    lp = -0.5 * diff**2 / variance - 0.5 * math.log(2 * math.pi * variance)
    return lp.sum(dim=-1)


def obs_indices(obs_times: Tensor, dt: float, n_steps: int) -> Tensor:
    """evidence_lower_bound.py:52: clamp(round(times / dt), max=T)."""
    return torch.clamp(torch.round(obs_times / dt).long(), max=n_steps)


@dataclass
class ElboTerms:
    """Per-trajectory [B] terms; the reference reports their batch means
    (evidence_lower_bound.py:62-74)."""

    obs: Tensor
    sde: Tensor
    gen: Tensor
    jac: Tensor

    def path_elbo(self) -> Tensor:
        """Path part of the ELBO (prior/posterior theta terms are O(B*P) PyTorch, out of scope)."""
        return (self.obs + self.sde - self.gen + self.jac).mean()


def elbo_terms(
    sde,
    z: Tensor,
    means: Tensor,
    chol: Tensor,
    theta: Tensor,
    dt: float,
    positive_dims: Sequence[int],
    obs_times: Tensor,
    obs_values: Tensor,
    obs_variance: float,
    obs_matrix: Tensor | None = None,
) -> ElboTerms:
    """The four path-dependent terms of ``compute_evidence_lower_bound``
    (evidence_lower_bound.py:28-56)."""
    B, T = z.shape[0], z.shape[1] - 1
    S = z.shape[-1]
    sqrt_dt = dt**0.5
    x = to_state(z, positive_dims)
    z_t, z_next = z[:, :-1], z[:, 1:]
    x_t, x_next = x[:, :-1], x[:, 1:]
    x_flat = x_t.reshape(B * T, S)
    th_flat = theta[:, None, :].expand(B, T, theta.shape[-1]).reshape(B * T, -1)
    drift = sde.drift(x_flat, th_flat).reshape(B, T, S)
    diffusion = sde.diffusion(x_flat, th_flat).reshape(B, T, S, S)
    sde_lp = gaussian_log_prob(x_next, x_t + drift * dt, diffusion * sqrt_dt)
    gen_lp = gaussian_log_prob(z_next, z_t + means * dt, chol * sqrt_dt)
    jac = log_jacobian(z, positive_dims)
    idx = obs_indices(obs_times, dt, T)
    obs_lp = gaussian_obs_log_prob(
        obs_values[None].expand(B, *obs_values.shape), x[:, idx], obs_variance, obs_matrix
    ).sum(dim=-1)
    return ElboTerms(obs_lp, sde_lp, gen_lp, jac)


# --------------------------------------------------------------------------------------
# synthetic inputs of SURVEY.md §8(d): one factory shared by tests, smoke and bench
# --------------------------------------------------------------------------------------
OU_OBS = (
    [0.0, 1.0, 2.0, 3.0, 4.0, 5.0],
    [[2.0], [1.5], [0.8], [1.2], [0.9], [1.1]],
)  # examples/ornstein_uhlenbeck.py:37-49
LV_OBS = (
    [0.0, 10.0, 20.0, 30.0, 40.0],
    [
        [71.0, 79.0],
        [47.61225908, 447.20971405],
        [80.53119269, 50.26254069],
        [23.10087379, 339.40432691],
        [158.05238324, 66.79611979],
    ],
)  # examples/lotka_volterra.py:53-64


@dataclass
class Problem:
    name: str
    sde: object
    weights: HeadWeights
    x0: Tensor  # latent z0 [B,S]
    context: Tensor  # [B,T,C] view of [B,T+1,C]
    theta: Tensor  # [B,P]
    eps: Tensor  # [B,T,S]
    dt: float
    positive_dims: tuple[int, ...]
    obs_times: Tensor
    obs_values: Tensor
    obs_variance: float


def make_problem(
    kind: str,
    batch: int,
    n_steps: int,
    *,
    dt: float = 0.05,
    context_dim: int = 256,
    hidden_dim: int = 64,
    num_layers: int = 2,
    state_dim: int | None = None,
    seed: int = 0,
    dtype: torch.dtype = torch.float32,
) -> Problem:
    g = torch.Generator().manual_seed(seed + 1)

    def randn(*shape: int) -> Tensor:
        return torch.randn(*shape, generator=g, dtype=torch.float64)

    T = n_steps
    if kind == "ou":
        sde, S, P = OrnsteinUhlenbeck(), 1, 3
        theta = torch.stack(
            [(0.3 * randn(batch)).exp(), 1.0 + 0.3 * randn(batch), (-1.0 + 0.3 * randn(batch)).exp()], -1
        )
        times, values, var = torch.tensor(OU_OBS[0]), torch.tensor(OU_OBS[1]), 0.1
    elif kind == "lv":
        sde, S, P = LotkaVolterra(), 2, 3
        theta = (torch.log(torch.tensor([0.5, 0.0025, 0.3], dtype=torch.float64)) + 0.1 * randn(batch, 3)).exp()
        times, values, var = torch.tensor(LV_OBS[0]), torch.tensor(LV_OBS[1]), 1.0
    elif kind == "l96":
        S = state_dim or 10
        sde, P = Lorenz96(S), 2
        theta = torch.stack([8.0 + 0.5 * randn(batch), (-1.0 + 0.2 * randn(batch)).exp()], -1)
        go = torch.Generator().manual_seed(1234)
        times = torch.arange(0.0, 5.01, 0.5)
        values = 8.0 + 3.0 * torch.randn(times.shape[0], S, generator=go)
        var = 0.25
    else:
        raise ValueError(kind)
    # keep only observations on the simulated horizon (examples use T_hor = last obs time)
    keep = times <= T * dt + 1e-9
    times, values = times[keep], values[keep]
    pos = tuple(sde.positive_dims)
    weights = make_head_weights(S, context_dim, P, hidden_dim, num_layers, seed=seed, dtype=dtype)
    ctx_full = (0.1 * randn(batch, T + 1, context_dim)).to(dtype)
    x0 = to_latent(values[0].to(torch.float64)[None].expand(batch, S).contiguous(), pos)
    return Problem(
        kind,
        sde,
        weights,
        x0.to(dtype),
        ctx_full[:, :-1],
        theta.to(dtype),
        randn(batch, T, S).to(dtype),
        dt,
        pos,
        times.to(dtype),
        values.to(dtype),
        var,
    )


def run_fwd_bwd(p: Problem, dtype: torch.dtype | None = None):
    """Reference CPU iteration: path fwd -> ELBO terms -> autograd backward of -path_elbo.

    Returns (paths, means, chol, terms, grads) where grads is a dict with x0, context,
    theta and the 4*NL+2 weight tensors in nn.GRU-native layout."""
    cast = (lambda t: t.to(dtype)) if dtype is not None else (lambda t: t)
    w = p.weights.map(lambda t: cast(t).detach().clone().requires_grad_(True))
    x0 = cast(p.x0).detach().clone().requires_grad_(True)
    ctx = cast(p.context).detach().clone().requires_grad_(True)
    theta = cast(p.theta).detach().clone().requires_grad_(True)
    eps = cast(p.eps)
    paths, means, chol = sample_paths(w, x0, ctx, theta, eps, p.dt)
    terms = elbo_terms(
        p.sde, paths, means, chol, theta, p.dt, p.positive_dims,
        cast(p.obs_times), cast(p.obs_values), p.obs_variance,
    )
    loss = -terms.path_elbo()
    loss.backward()
    grads = {"x0": x0.grad, "context": ctx.grad, "theta": theta.grad}
    for k in range(w.num_layers):
        grads[f"w_ih_l{k}"] = w.w_ih[k].grad
        grads[f"w_hh_l{k}"] = w.w_hh[k].grad
        grads[f"b_ih_l{k}"] = w.b_ih[k].grad
        grads[f"b_hh_l{k}"] = w.b_hh[k].grad
    grads["out_w"] = w.out_w.grad
    grads["out_b"] = w.out_b.grad
    return paths.detach(), means.detach(), chol.detach(), terms, grads


# --------------------------------------------------------------------------------------
# Callers either side of the path: prior simulator, posterior summary, optimiser tail
# --------------------------------------------------------------------------------------
def euler_maruyama(sde, x0: Tensor, theta: Tensor, n_steps: int, dt: float, positive_dims: Sequence[int] = (),
                   noise: Tensor | None = None) -> Tensor:
    """core/euler_maruyama.py:27-45: explicit Euler-Maruyama with a clamp of the positive dims after every step."""
    sqrt_dt = dt**0.5
    traj = [x0]
    x = x0
    pos = list(positive_dims)
    for step in range(n_steps):
        x = x + sde.drift(x, theta) * dt + torch.einsum("bij,bj->bi", sde.diffusion(x, theta), noise[:, step]) * sqrt_dt
        if pos:
            keep = torch.ones(x.shape[-1], dtype=torch.bool)
            keep[pos] = False
            x = torch.where(keep, x, x.clamp(min=1e-6))  # out-of-place form of x[:, pos] = x[:, pos].clamp(min=1e-6)
        traj.append(x)
    return torch.stack(traj, dim=1)


def pretrain_mse(sde, theta: Tensor, obs_times: Tensor, obs_values: Tensor, n_steps: int, dt: float,
                 positive_dims: Sequence[int], noise: Tensor) -> Tensor:
    """inference/trainer.py:253-259."""
    n = theta.shape[0]
    x0 = obs_values[0].unsqueeze(0).expand(n, -1).to(theta.dtype)
    paths = euler_maruyama(sde, x0, theta, n_steps, dt, positive_dims, noise)
    obs_idx = (obs_times / dt).round().long()
    return ((paths[:, obs_idx] - obs_values.to(theta.dtype)) ** 2).mean()


def philox_normal(seed: int, B: int, T: int, S: int, group: int = 0):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011: multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments
    0x9E3779B9 / 0xBB67AE85) with counter (t_lo, t_hi, b_lo, b_hi) and key = seed, then Box-Muller on
    u = (top 24 bits + 0.5) 2^-24: the in-kernel noise of csrc/em.cu, restated in numpy (float64 transcendental part)."""
    import numpy as np

    if S > 4:  # wide state spaces: one Philox block per four state dims, the group index in bits 28.. of the second counter word
        return torch.cat([philox_normal(seed, B, T, min(4, S - 4 * g), group=g) for g in range((S + 3) // 4)], dim=-1)
    b, t = np.meshgrid(np.arange(B, dtype=np.uint64), np.arange(T, dtype=np.uint64), indexing="ij")
    m32 = np.uint64(0xFFFFFFFF)
    c = [t & m32, (t >> np.uint64(32)) | np.uint64(group << 28), b & m32, b >> np.uint64(32)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]  # 32 x 32 -> 64-bit products (no overflow in uint64)
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    u = [((x >> np.uint64(8)).astype(np.float64) + 0.5) * 2.0**-24 for x in c]
    ra, rb = np.sqrt(-2.0 * np.log(u[0])), np.sqrt(-2.0 * np.log(u[2]))
    n = np.stack([ra * np.cos(2 * np.pi * u[1]), ra * np.sin(2 * np.pi * u[1]),
                  rb * np.cos(2 * np.pi * u[3]), rb * np.sin(2 * np.pi * u[3])], axis=-1)
    return torch.from_numpy(n[..., :S].copy())


def path_summary(z: Tensor, positive_dims: Sequence[int]):
    """posterior/variational_posterior.py:116-135: x = from_latent(z); mean / std (Bessel) over the samples."""
    x = to_state(z, positive_dims)
    return x, x.mean(dim=0), x.std(dim=0)


def adamw_ema_steps(params: list[Tensor], grads_per_step: list[list[Tensor]], lrs_by_param: list[float], max_norm: float,
                    ema_decay: float, inv_scale: float = 1.0):
    """inference/trainer.py:199-203 + :126 with the third-party pieces they call: ``GradScaler.unscale_`` (grad *= 1/scale),
    ``nn.utils.clip_grad_norm_``, ``torch.optim.AdamW`` (torch defaults, one param group per learning rate) and
    ``ExponentialMovingAverage.update`` (shadow.lerp_(param, 1 - decay), exponential_moving_average.py:25-28)."""
    ps = [torch.nn.Parameter(p.clone()) for p in params]
    opt = torch.optim.AdamW([{"params": [p], "lr": lr} for p, lr in zip(ps, lrs_by_param)])
    shadow = [p.detach().clone() for p in ps]
    norms = []
    for grads in grads_per_step:
        for p, g in zip(ps, grads):
            p.grad = g.clone() * inv_scale
        norms.append(torch.nn.utils.clip_grad_norm_(ps, max_norm) if max_norm > 0 else
                     torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(p.grad) for p in ps])))
        opt.step()
        with torch.no_grad():
            for s, p in zip(shadow, ps):
                s.lerp_(p.detach(), 1 - ema_decay)
    return [p.detach() for p in ps], shadow, torch.stack(norms)
