"""Observation value types with the reference's names and validation
(src/variational_sde/core/observations.py:12-74), as plain dataclasses."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Protocol, runtime_checkable

import torch
from torch import Tensor


@dataclass(frozen=True)
class Observations:
    times: Tensor
    values: Tensor

    def __post_init__(self) -> None:
        if self.times.ndim != 1:
            raise ValueError("times must be 1D tensor")
        if self.values.ndim != 2:
            raise ValueError("values must be 2D tensor [T_obs, obs_dim]")
        if self.times.shape[0] != self.values.shape[0]:
            raise ValueError(
                f"times and values must have same first dimension: got {self.times.shape[0]} vs {self.values.shape[0]}")
        if not torch.all(self.times[1:] >= self.times[:-1]):
            raise ValueError("times must be sorted in non-decreasing order")


@runtime_checkable
class ObservationLikelihood(Protocol):
    def log_prob(self, observations: Tensor, state: Tensor) -> Tensor: ...


@dataclass(frozen=True)
class GaussianObservationLikelihood:
    variance: float
    obs_matrix: Optional[Tensor] = None

    def __post_init__(self) -> None:
        if self.variance <= 0:
            raise ValueError("variance must be positive")

    def log_prob(self, observations: Tensor, state: Tensor) -> Tensor:
        if self.obs_matrix is not None:
            if self.obs_matrix.ndim != 2:
                raise ValueError("obs_matrix must be 2D [obs_dim, state_dim]")
            if self.obs_matrix.shape[0] != observations.shape[-1]:
                raise ValueError("obs_matrix first dim must match observations")
            if self.obs_matrix.shape[1] != state.shape[-1]:
                raise ValueError("obs_matrix second dim must match state")
            predicted = torch.einsum("od,...d->...o", self.obs_matrix, state)
        else:
            predicted = state
        if observations.shape != predicted.shape:
            raise ValueError(
                f"observation shape {observations.shape} does not match predicted shape {predicted.shape}")
        diff = observations - predicted
        log_prob = -0.5 * (diff**2) / self.variance - 0.5 * math.log(2 * math.pi * self.variance)
        return log_prob.sum(dim=-1)
