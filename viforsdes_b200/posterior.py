"""Sampling side of ``VariationalPosterior`` (src/variational_sde/posterior/variational_posterior.py:93-135): draw n
posterior trajectories with the stash-less forward and summarise them.

``summarise_paths`` is one fused pass (``visde_path_summary``, csrc/summary.cu) over the latent paths: the softplus
state-space map (``DiffusionPathSample.x``), the per-grid-point mean and the Bessel-corrected standard deviation, where the
reference makes three passes (``to_state`` clone + index_put, ``mean(dim=0)``, ``std(dim=0)``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from viforsdes_b200 import _lib
from viforsdes_b200.observations import Observations
from viforsdes_b200.ops import _f32c, _ptr, _require_cuda, _stream
from viforsdes_b200.sampler import EncoderProtocol, HeadProtocol, sample_diffusion_paths
from viforsdes_b200.state_space import StateSpace

QUANTILE_LEVELS = (0.05, 0.25, 0.5, 0.75, 0.95)  # variational_posterior.py:24


@dataclass(frozen=True)
class VariationalPosteriorSamples:
    sde_parameters: Tensor
    diffusion_paths: Tensor


@dataclass(frozen=True)
class Quantiles:
    q05: Tensor
    q25: Tensor
    q50: Tensor
    q75: Tensor
    q95: Tensor


@dataclass(frozen=True)
class VariationalPosteriorSummary:
    sde_parameter_mean: Tensor
    sde_parameter_std: Tensor
    sde_parameter_quantiles: Quantiles
    diffusion_path_mean: Tensor
    diffusion_path_std: Tensor


def summarise_paths(z: Tensor, state_space: StateSpace, want_x: bool = True) -> tuple[Optional[Tensor], Tensor, Tensor]:
    """z [n, T+1, S] latent -> (x = from_latent(z) or None, mean [T+1, S], std [T+1, S]) over the n samples."""
    _require_cuda(z)
    if z.dim() != 3 or z.shape[2] != state_space.dim:
        raise ValueError(f"z must be [n, T+1, {state_space.dim}], got {tuple(z.shape)}")
    lib = _lib.load()
    n, T1, S = z.shape
    zf = _f32c(z)
    f = dict(device=z.device, dtype=torch.float32)
    x = torch.empty(n, T1, S, **f) if want_x else None
    mean, std = torch.empty(T1, S, **f), torch.empty(T1, S, **f)
    ws_bytes = lib.visde_path_summary_workspace_bytes(n, T1, S)
    ws = torch.empty(ws_bytes, device=z.device, dtype=torch.uint8)
    with torch.cuda.device(z.device):
        _lib.check(lib.visde_path_summary(n, T1, S, state_space.positive_mask, _ptr(zf), _ptr(x), _ptr(mean), _ptr(std),
                                          _ptr(ws), ws_bytes, _stream()))
    return x, mean, std


class _EvalMode:
    """``self.model.eval()`` of variational_posterior.py:96 for the modules this function is handed: the head AND the
    encoder (dropout / other train-mode behaviour must be off while sampling); restores the previous modes on exit."""

    def __init__(self, *modules) -> None:
        self.modules = [m for m in modules if isinstance(m, torch.nn.Module)]

    def __enter__(self):
        self.was = [m.training for m in self.modules]
        for m in self.modules:
            m.eval()
        return self

    def __exit__(self, *exc) -> None:
        for m, w in zip(self.modules, self.was):
            m.train(w)


@torch.no_grad()
def sample_posterior(encoder: EncoderProtocol, head: HeadProtocol, sde_parameter_posterior, observations: Observations,
                     n: int, time_horizon: float, time_step: float, state_space: StateSpace,
                     noise: Optional[Tensor] = None) -> VariationalPosteriorSamples:
    """``VariationalPosterior.sample`` (variational_posterior.py:93-114) without the EMA swap (the caller owns it)."""
    with _EvalMode(encoder, head):
        sde_parameters = sde_parameter_posterior.rsample(n)
        x0 = observations.values[0].unsqueeze(0).expand(n, -1).contiguous()
        result = sample_diffusion_paths(encoder, head, observations, sde_parameters, x0, time_horizon, time_step,
                                        state_space, noise=noise)
        x, _, _ = summarise_paths(result.z, state_space) if result.z.is_cuda else (result.x, None, None)
    return VariationalPosteriorSamples(sde_parameters=sde_parameters, diffusion_paths=x)


@torch.no_grad()
def summarise_posterior(encoder: EncoderProtocol, head: HeadProtocol, sde_parameter_posterior, observations: Observations,
                        time_horizon: float, time_step: float, state_space: StateSpace, n_samples: int = 1000,
                        noise: Optional[Tensor] = None) -> VariationalPosteriorSummary:
    """``VariationalPosterior.summary`` (variational_posterior.py:116-135)."""
    sde_parameters = sde_parameter_posterior.rsample(n_samples)
    x0 = observations.values[0].unsqueeze(0).expand(n_samples, -1).contiguous()
    with _EvalMode(encoder, head):
        result = sample_diffusion_paths(encoder, head, observations, sde_parameters, x0, time_horizon, time_step,
                                        state_space, noise=noise)
    _, mean, std = summarise_paths(result.z, state_space, want_x=False)
    q = torch.quantile(sde_parameters, torch.tensor(QUANTILE_LEVELS, device=sde_parameters.device,
                                                    dtype=sde_parameters.dtype), dim=0)
    return VariationalPosteriorSummary(
        sde_parameter_mean=sde_parameters.mean(dim=0), sde_parameter_std=sde_parameters.std(dim=0),
        sde_parameter_quantiles=Quantiles(q05=q[0], q25=q[1], q50=q[2], q75=q[3], q95=q[4]),
        diffusion_path_mean=mean, diffusion_path_std=std)
