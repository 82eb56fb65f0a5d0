"""SDE protocol (src/variational_sde/core/sde.py:8-48, kept verbatim in shape) plus the two
built-in models whose drift/diffusion also exist as device functors in csrc/elbo.cu."""
from __future__ import annotations

from typing import Callable, Protocol, runtime_checkable

import torch
from torch import Tensor

from viforsdes_b200 import _lib


@runtime_checkable
class SDE(Protocol):
    state_dim: int
    sde_param_dim: int

    def drift(self, x: Tensor, sde_parameters: Tensor) -> Tensor: ...
    def diffusion(self, x: Tensor, sde_parameters: Tensor) -> Tensor: ...


class FunctionalSDE:
    def __init__(self, drift_fn: Callable[[Tensor, Tensor], Tensor], diffusion_fn: Callable[[Tensor, Tensor], Tensor],
                 state_dim: int, sde_param_dim: int) -> None:
        self._drift_fn = drift_fn
        self._diffusion_fn = diffusion_fn
        self.state_dim = state_dim
        self.sde_param_dim = sde_param_dim

    def drift(self, x: Tensor, sde_parameters: Tensor) -> Tensor:
        return self._drift_fn(x, sde_parameters)

    def diffusion(self, x: Tensor, sde_parameters: Tensor) -> Tensor:
        return self._diffusion_fn(x, sde_parameters)


def make_sde(drift: Callable[[Tensor, Tensor], Tensor], diffusion: Callable[[Tensor, Tensor], Tensor],
             state_dim: int, sde_param_dim: int) -> SDE:
    return FunctionalSDE(drift, diffusion, state_dim, sde_param_dim)


class OrnsteinUhlenbeck:
    """examples/ornstein_uhlenbeck.py:18-30; `device_kind` routes the ELBO to the fused functor."""

    state_dim = 1
    sde_param_dim = 3
    device_kind = _lib.SDE_OU

    def drift(self, x: Tensor, sde_parameters: Tensor) -> Tensor:
        return sde_parameters[..., 0:1] * (sde_parameters[..., 1:2] - x)

    def diffusion(self, x: Tensor, sde_parameters: Tensor) -> Tensor:
        return sde_parameters[..., 2:3].reshape(x.shape[0], 1, 1)


class LotkaVolterra:
    """examples/lotka_volterra.py:18-46."""

    state_dim = 2
    sde_param_dim = 3
    device_kind = _lib.SDE_LV

    def drift(self, x: Tensor, sde_parameters: Tensor) -> Tensor:
        u, v = x[..., 0], x[..., 1]
        t1, t2, t3 = sde_parameters[..., 0], sde_parameters[..., 1], sde_parameters[..., 2]
        return torch.stack([t1 * u - t2 * u * v, t2 * u * v - t3 * v], dim=-1)

    def diffusion(self, x: Tensor, sde_parameters: Tensor) -> Tensor:
        u, v = x[..., 0], x[..., 1]
        t1, t2, t3 = sde_parameters[..., 0], sde_parameters[..., 1], sde_parameters[..., 2]
        uv = u * v
        b11, b12, b22 = t1 * u + t2 * uv, -t2 * uv, t3 * v + t2 * uv
        L00 = torch.sqrt(b11.clamp(min=1e-6))
        L10 = b12 / L00.clamp(min=1e-6)
        L11 = torch.sqrt((b22 - L10**2).clamp(min=1e-6))
        zeros = torch.zeros_like(L00)
        return torch.stack([torch.stack([L00, zeros], -1), torch.stack([L10, L11], -1)], -2)
