"""``compute_evidence_lower_bound`` with the signature and result types of
src/variational_sde/inference/evidence_lower_bound.py:19-74.

The four path-dependent terms come from the time-parallel kernel ``visde::elbo_fwd``
(csrc/elbo.cu): built-in OU / Lotka-Volterra models (objects carrying ``device_kind``) are
evaluated entirely on-device; any other ``SDE`` has its ``drift`` / ``diffusion`` callbacks run in
PyTorch on the flattened [B*T, S] path exactly as the reference does (:37-40) and only the
Gaussian algebra is fused.  Prior and theta-posterior terms are O(B*P) and stay in PyTorch.
"""
from __future__ import annotations

import torch
from torch import Tensor

from viforsdes_b200 import _lib, ops  # noqa: F401
from viforsdes_b200.observations import GaussianObservationLikelihood, ObservationLikelihood, Observations
from viforsdes_b200.types import DiffusionPathSample, EvidenceLowerBoundComponents, EvidenceLowerBoundResult


def observation_indices(times: Tensor, time_step: float, n_steps: int) -> Tensor:
    """evidence_lower_bound.py:52."""
    return torch.clamp(torch.round(times / time_step).long(), max=n_steps)


def path_elbo_terms(sde, observations: Observations, observation_likelihood: ObservationLikelihood,
                    sde_parameters: Tensor, sample: DiffusionPathSample, time_step: float) -> Tensor:
    """[B,4] per-trajectory (obs, sde, gen, jacobian) log-probabilities, differentiable."""
    z = sample.z
    B, n_steps = z.shape[0], z.shape[1] - 1
    S = z.shape[2]
    mask = sample.state_space.positive_mask
    obs_idx = observation_indices(observations.times.to(z.device), time_step, n_steps)
    gaussian = isinstance(observation_likelihood, GaussianObservationLikelihood)
    if gaussian:
        k_idx, k_vals = obs_idx, observations.values.to(z.device)
        k_mat, k_var = observation_likelihood.obs_matrix, observation_likelihood.variance
        if k_mat is not None:
            k_mat = k_mat.to(z.device)
    else:
        k_idx = torch.zeros(0, dtype=torch.long, device=z.device)
        k_vals = torch.zeros(0, S, device=z.device)
        k_mat, k_var = None, 1.0
    kind = getattr(sde, "device_kind", _lib.SDE_GENERIC)
    drift = diffusion = None
    if kind == _lib.SDE_GENERIC:
        x_t = sample.x[:, :-1]
        x_flat = x_t.reshape(B * n_steps, S)
        th_flat = sde_parameters[:, None, :].expand(B, n_steps, sde_parameters.shape[-1]).reshape(B * n_steps, -1)
        drift = sde.drift(x_flat, th_flat).reshape(B, n_steps, S)
        diffusion = sde.diffusion(x_flat, th_flat).reshape(B, n_steps, S, S)
    terms = torch.ops.visde.elbo_fwd(z, sample.transition_means, sample.transition_cholesky, sde_parameters, drift,
                                     diffusion, k_idx, k_vals, k_mat, float(k_var), float(time_step), int(kind),
                                     int(mask))
    if not gaussian:
        x = sample.x
        obs_lp = observation_likelihood.log_prob(
            observations.values.to(z.device)[None].expand(B, *observations.values.shape), x[:, obs_idx]).sum(dim=-1)
        terms = torch.cat([obs_lp[:, None].to(terms.dtype), terms[:, 1:]], dim=1)
    return terms


def compute_evidence_lower_bound(sde, observations: Observations, observation_likelihood: ObservationLikelihood,
                                 prior, sde_parameter_posterior, sde_parameters: Tensor, sample: DiffusionPathSample,
                                 time_step: float) -> EvidenceLowerBoundResult:
    terms = path_elbo_terms(sde, observations, observation_likelihood, sde_parameters, sample, time_step)
    obs_lp, sde_lp, gen_lp, jac = terms[:, 0], terms[:, 1], terms[:, 2], terms[:, 3]
    prior_lp = prior.log_prob(sde_parameters)
    if prior_lp.ndim > 1:
        prior_lp = prior_lp.sum(-1)
    post_lp = sde_parameter_posterior.log_prob(sde_parameters)
    elbo = obs_lp + sde_lp - gen_lp + jac + prior_lp - post_lp  # evidence_lower_bound.py:64
    return EvidenceLowerBoundResult(
        evidence_lower_bound=elbo.mean(),
        components=EvidenceLowerBoundComponents(
            observation_log_prob=obs_lp.mean(), sde_log_prob=sde_lp.mean(), generative_log_prob=gen_lp.mean(),
            prior_log_prob=prior_lp.mean(), posterior_log_prob=post_lp.mean()))
