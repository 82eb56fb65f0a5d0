"""Softplus state-space transform with the interface of src/variational_sde/inference/state_space.py:8-38
(``to_state`` / ``to_latent`` / ``log_jacobian``, ``dim``, ``positive_dims``).  Host-side helper: on the hot path the
transform and its log-Jacobian are evaluated inside the ELBO kernel (csrc/elbo.cu) from ``positive_mask``.

Implemented with a boolean column mask and ``torch.where`` instead of index assignment into a clone: one fused
elementwise pass, no gather / scatter kernels, and ``positive_mask`` (bit s = dim s is positive) is the very word
the kernels take."""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor
from torch.nn import functional as F


class StateSpace:
    def __init__(self, dim: int, positive_dims: Sequence[int] | None = None) -> None:
        dims = [int(d) for d in (positive_dims or ())]
        if dim < 1:
            raise ValueError(f"dim must be >= 1, got {dim}")
        if not all(0 <= d < dim for d in dims):
            raise ValueError(f"positive_dims must be in [0, {dim}), got {dims}")
        if len(set(dims)) != len(dims):
            raise ValueError(f"positive_dims must be unique, got {dims}")
        self.dim = dim
        self.positive_dims = dims
        self.positive_mask = sum(1 << d for d in dims)
        self._cols: dict = {}  # device -> bool [dim]

    def _positive_columns(self, like: Tensor) -> Tensor:
        cols = self._cols.get(like.device)
        if cols is None:
            cols = torch.zeros(self.dim, dtype=torch.bool, device=like.device)
            cols[self.positive_dims] = True
            self._cols[like.device] = cols
        return cols

    def to_state(self, z: Tensor) -> Tensor:
        """x = softplus(z) on the positive dims (torch default beta = 1, threshold = 20), identity elsewhere."""
        if self.positive_mask == 0:
            return z
        return torch.where(self._positive_columns(z), F.softplus(z), z)

    def to_latent(self, x: Tensor) -> Tensor:
        """Inverse of ``to_state``: z = x + log(1 - exp(-x)) with x floored at 1e-6 on the positive dims."""
        if self.positive_mask == 0:
            return x
        xp = x.clamp(min=1e-6)
        return torch.where(self._positive_columns(x), xp + torch.log(-torch.expm1(-xp)), x)

    def log_jacobian(self, z: Tensor) -> Tensor:
        """sum over positive dims of log d softplus(z) / dz = logsigmoid(z); shape z.shape[:-1]."""
        if self.positive_mask == 0:
            return z.new_zeros(z.shape[:-1])
        cols = self._positive_columns(z)
        return torch.where(cols, F.logsigmoid(z), torch.zeros_like(z)).sum(dim=-1)
