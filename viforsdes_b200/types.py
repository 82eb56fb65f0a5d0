"""Result types with the reference's field names (src/variational_sde/inference/types.py:12-45)."""
from __future__ import annotations

from dataclasses import dataclass

from torch import Tensor

from viforsdes_b200.state_space import StateSpace


@dataclass(frozen=True)
class DiffusionPathSample:
    z: Tensor
    transition_means: Tensor
    transition_cholesky: Tensor
    state_space: StateSpace

    @property
    def x(self) -> Tensor:
        return self.state_space.to_state(self.z)

    def log_jacobian(self) -> Tensor:
        return self.state_space.log_jacobian(self.z[:, 1:]).sum(dim=-1)


@dataclass(frozen=True)
class EvidenceLowerBoundComponents:
    observation_log_prob: Tensor
    sde_log_prob: Tensor
    generative_log_prob: Tensor
    prior_log_prob: Tensor
    posterior_log_prob: Tensor


@dataclass(frozen=True)
class EvidenceLowerBoundResult:
    evidence_lower_bound: Tensor
    components: EvidenceLowerBoundComponents
