"""``sample_diffusion_paths`` with the signature of
src/variational_sde/inference/diffusion_path_sampler.py:35-69, plus an optional ``noise``
argument so that tests / benchmarks can inject fixed standard-normal draws."""
from __future__ import annotations

from typing import Optional, Protocol

import torch
from torch import Tensor

from viforsdes_b200.observations import Observations
from viforsdes_b200.state_space import StateSpace
from viforsdes_b200.types import DiffusionPathSample


class EncoderProtocol(Protocol):
    def __call__(self, obs_values: Tensor, obs_times: Tensor, sde_parameters: Tensor, time_horizon: float,
                 time_step: float) -> Tensor: ...


class HeadProtocol(Protocol):
    def sample_diffusion_paths(self, x0: Tensor, context: Tensor, sde_parameters: Tensor, standard_noise: Tensor,
                               time_step: float) -> tuple[Tensor, Tensor, Tensor]: ...


def sample_diffusion_paths(encoder: EncoderProtocol, head: HeadProtocol, observations: Observations,
                           sde_parameters: Tensor, x0: Tensor, time_horizon: float, time_step: float,
                           state_space: StateSpace, noise: Optional[Tensor] = None, seed: Optional[int] = None,
                           batch_offset: int = 0) -> DiffusionPathSample:
    """`noise`: injected standard normals (tests, benchmarks).  `seed`: opt-in counter-based noise instead of the reference's
    ``torch.randn`` (diffusion_path_sampler.py:57): the library's Philox4x32-10 stream, a pure function of
    (seed, trajectory index, grid step, state dim) -- reproducible run to run and independent of the rank layout when each
    rank passes its shard's first trajectory index as `batch_offset`.  Neither: ``torch.randn`` as in the reference."""
    batch_size, state_dim = x0.shape
    context = encoder(observations.values, observations.times, sde_parameters, time_horizon, time_step)
    n_steps = context.shape[1] - 1
    if noise is None and seed is not None and x0.is_cuda:
        from viforsdes_b200.euler_maruyama import philox_normal

        noise = philox_normal(int(seed), batch_offset + batch_size, n_steps, state_dim, device=x0.device)[batch_offset:].to(x0.dtype)
    elif noise is None:
        noise = torch.randn(batch_size, n_steps, state_dim, device=x0.device, dtype=x0.dtype)
    z0 = state_space.to_latent(x0)
    # the context at the final grid time is dropped (diffusion_path_sampler.py:61); the kernels read
    # the strided view in place
    paths, means, chol = head.sample_diffusion_paths(z0, context[:, :-1], sde_parameters, noise, time_step)
    return DiffusionPathSample(z=paths, transition_means=means, transition_cholesky=chol, state_space=state_space)
