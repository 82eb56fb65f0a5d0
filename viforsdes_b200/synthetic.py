"""Synthetic inputs of the benchmark workloads: seeded on the CPU so every arm
(CUDA, host session, CPU baseline) sees identical tensors.  tests/test_synthetic.py keeps this in
lock-step with the oracle's own factory."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import torch
from torch import Tensor

from viforsdes_b200 import _lib

OU_OBS = ([0.0, 1.0, 2.0, 3.0, 4.0, 5.0], [[2.0], [1.5], [0.8], [1.2], [0.9], [1.1]])
LV_OBS = ([0.0, 10.0, 20.0, 30.0, 40.0],
          [[71.0, 79.0], [47.61225908, 447.20971405], [80.53119269, 50.26254069], [23.10087379, 339.40432691],
           [158.05238324, 66.79611979]])

# name -> (kind, B, T, dt): the configurations of BASELINE.json
WORKLOADS = {
    "ou_b128_t100": ("ou", 128, 100, 0.05),      # configs[0]
    "lv_b128_t800": ("lv", 128, 800, 0.05),      # configs[1]  (bench default at N=1)
    "ou_b65536_t100": ("ou", 65536, 100, 0.05),  # configs[2]  top of the batch sweep
    "ou_b8192_t100": ("ou", 8192, 100, 0.05),
    "lv_b128_t20000": ("lv", 128, 20000, 0.002),  # configs[3]
    "l96_b8192_t100": ("l96", 8192, 100, 0.05),  # configs[4]
}


class _Workloads(dict):
    """Named configurations plus the pattern `<ou|lv|l96>_b<B>_t<T>` (dt = 0.05) for batch sweeps."""

    def __missing__(self, name: str):
        import re

        m = re.fullmatch(r"(ou|lv|l96)_b(\d+)_t(\d+)", name)
        if not m:
            raise KeyError(name)
        return (m.group(1), int(m.group(2)), int(m.group(3)), 0.05)


WORKLOADS = _Workloads(WORKLOADS)


class Lorenz96:
    """The user-defined SDE of BASELINE.json configs[4], written against the reference's ``SDE``
    protocol (src/variational_sde/core/sde.py:8-15) and evaluated in PyTorch on the flattened path exactly as
    inference/evidence_lower_bound.py:37-40 calls it: cyclic ``drift_i = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + F``,
    ``diffusion = sigma I``; sde_parameters = (F, sigma)."""

    sde_param_dim = 2

    def __init__(self, state_dim: int = 10) -> None:
        self.state_dim = state_dim

    def drift(self, x: Tensor, sde_parameters: Tensor) -> Tensor:
        ahead, behind, behind2 = (torch.roll(x, k, dims=-1) for k in (-1, 1, 2))
        return (ahead - behind2) * behind - x + sde_parameters[..., 0:1]

    def diffusion(self, x: Tensor, sde_parameters: Tensor) -> Tensor:
        # sigma I written as diag_embed: the same values as ``sigma[..., None, None] * eye`` (the oracle's form), but its autograd
        # backward reads the S diagonal entries of the cotangent instead of multiplying and reducing the whole [.., S, S] block
        return torch.diag_embed(sde_parameters[..., 1:2].expand(*x.shape[:-1], self.state_dim))


@dataclass
class Inputs:
    kind: str
    sde_kind: int
    positive_dims: Tuple[int, ...]
    x0: Tensor
    context_full: Tensor  # [B,T+1,C]; the head reads context_full[:, :-1]
    theta: Tensor
    eps: Tensor
    w_ih: List[Tensor]
    w_hh: List[Tensor]
    b_ih: List[Tensor]
    b_hh: List[Tensor]
    out_w: Tensor
    out_b: Tensor
    dt: float
    obs_times: Tensor
    obs_values: Tensor
    obs_variance: float
    sde: object = None  # user SDE object (sde_kind GENERIC): drift / diffusion evaluated in PyTorch

    @property
    def positive_mask(self) -> int:
        m = 0
        for d in self.positive_dims:
            m |= 1 << d
        return m

    @property
    def obs_idx(self) -> Tensor:
        T = self.context_full.shape[1] - 1
        return torch.clamp(torch.round(self.obs_times / self.dt).long(), max=T)


def _softplus_inverse(x: Tensor) -> Tensor:
    x = x.clamp(min=1e-6)
    return x + torch.log(-torch.expm1(-x))


def make_inputs(kind: str, batch: int, n_steps: int, *, dt: float = 0.05, context_dim: int = 256,
                hidden_dim: int = 64, num_layers: int = 2, state_dim: int | None = None, seed: int = 0) -> Inputs:
    g = torch.Generator().manual_seed(seed)
    H = hidden_dim
    k = 1.0 / math.sqrt(H)

    def u(*shape: int) -> Tensor:
        return ((torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1) * k).to(torch.float32)

    if kind == "ou":
        S, P, pos = 1, 3, ()
    elif kind == "lv":
        S, P, pos = 2, 3, (0, 1)
    elif kind == "l96":
        S, P, pos = state_dim or 10, 2, ()
    else:
        raise ValueError(kind)
    w_ih = [u(3 * H, S + context_dim + P)] + [u(3 * H, H) for _ in range(num_layers - 1)]
    w_hh = [u(3 * H, H) for _ in range(num_layers)]
    b_ih = [u(3 * H) for _ in range(num_layers)]
    b_hh = [u(3 * H) for _ in range(num_layers)]
    n_tril = S * (S + 1) // 2
    out_w = (torch.randn(S + n_tril, H, generator=g, dtype=torch.float64) * 0.1).to(torch.float32)
    out_b = torch.zeros(S + n_tril)
    for d in range(S):
        out_b[S + d * (d + 3) // 2] = 1.0

    g2 = torch.Generator().manual_seed(seed + 1)

    def randn(*shape: int) -> Tensor:
        return torch.randn(*shape, generator=g2, dtype=torch.float64)

    T = n_steps
    if kind == "ou":
        theta = torch.stack([(0.3 * randn(batch)).exp(), 1.0 + 0.3 * randn(batch), (-1.0 + 0.3 * randn(batch)).exp()], -1)
        times, values, var, sk = torch.tensor(OU_OBS[0]), torch.tensor(OU_OBS[1]), 0.1, _lib.SDE_OU
    elif kind == "lv":
        theta = (torch.log(torch.tensor([0.5, 0.0025, 0.3], dtype=torch.float64)) + 0.1 * randn(batch, 3)).exp()
        times, values, var, sk = torch.tensor(LV_OBS[0]), torch.tensor(LV_OBS[1]), 1.0, _lib.SDE_LV
    else:
        theta = torch.stack([8.0 + 0.5 * randn(batch), (-1.0 + 0.2 * randn(batch)).exp()], -1)
        go = torch.Generator().manual_seed(1234)
        times = torch.arange(0.0, 5.01, 0.5)
        values = 8.0 + 3.0 * torch.randn(times.shape[0], S, generator=go)
        var, sk = 0.25, _lib.SDE_GENERIC
    keep = times <= T * dt + 1e-9
    times, values = times[keep], values[keep]
    ctx_full = (0.1 * randn(batch, T + 1, context_dim)).to(torch.float32)
    x0 = values[0].to(torch.float64)[None].expand(batch, S).contiguous()
    if pos:
        x0 = x0.clone()
        x0[:, list(pos)] = _softplus_inverse(x0[:, list(pos)])
    eps = randn(batch, T, S).to(torch.float32)
    return Inputs(kind, sk, pos, x0.to(torch.float32), ctx_full, theta.to(torch.float32), eps, w_ih, w_hh, b_ih, b_hh,
                  out_w, out_b, dt, times.to(torch.float32), values.to(torch.float32), var,
                  Lorenz96(S) if kind == "l96" else None)
