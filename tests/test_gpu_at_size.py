"""Parity at the BASELINE.json sizes (VERDICT r1 "close the parity gaps at size"): every configuration is run on the
GPU at its own shape and checked against the CPU oracle -- completely where the oracle finishes in seconds
(configs 1, 2), and on batch slices (per-trajectory outputs and gradients are independent across trajectories)
plus a small complete batch for the weight gradients where it does not (configs 3, 4, 5)."""
from __future__ import annotations

import dataclasses

import pytest
import torch

from oracle import oracle_torch as O
from tests._util import assert_parity, check_iteration, oracle_refs, run_cuda_fwd_bwd

pytestmark = pytest.mark.gpu
FULL = dict(context_dim=256, hidden_dim=64, num_layers=2)
PER_TRAJECTORY = ("x0", "context", "theta")


def _slice(p: O.Problem, rows) -> O.Problem:
    return dataclasses.replace(p, x0=p.x0[rows], context=p.context[rows], theta=p.theta[rows], eps=p.eps[rows])


def _check_slice(cuda_out, p: O.Problem, rows, tag: str):
    """Forward outputs, ELBO terms and the per-trajectory gradients of the rows `rows` against the oracle run on just
    those rows (the loss is a batch mean: gradients of a slice of n rows of a batch of B scale by n / B)."""
    paths, means, chol, terms, grads = cuda_out
    sub = _slice(p, rows)
    r32, r64 = oracle_refs(sub)
    n, B = sub.x0.shape[0], p.x0.shape[0]
    for a, b32, b64, nm in zip((paths, means, chol), r32[:3], r64[:3], ("paths", "means", "chol")):
        assert_parity(a[rows], b32, b64, name=f"{tag}{nm}")
    tscale = max(getattr(r64[3], nm).abs().max().item() for nm in ("obs", "sde", "gen", "jac"))
    for j, nm in enumerate(("obs", "sde", "gen", "jac")):  # atol_abs: see tests/_util.check_iteration
        assert_parity(terms[rows, j], getattr(r32[3], nm), getattr(r64[3], nm), name=f"{tag}term_{nm}", atol_abs=1e-7 * tscale)
    for nm in PER_TRAJECTORY:
        assert_parity(grads[nm][rows] * (B / n), r32[4][nm], r64[4][nm], name=f"{tag}grad_{nm}")


def test_config1_ou_b128_t100_full():
    """BASELINE configs[0] at size: every output, term and gradient against the fp32 + fp64 oracle."""
    p = O.make_problem("ou", 128, 100, **FULL)
    r32, r64 = oracle_refs(p)
    check_iteration(run_cuda_fwd_bwd(p), r32, r64, tag="cfg1/")


def test_config2_lv_b128_t800_full():
    """BASELINE configs[1] at size (the oracle takes ~20 s for the fp32 + fp64 pair)."""
    p = O.make_problem("lv", 128, 800, **FULL)
    r32, r64 = oracle_refs(p)
    check_iteration(run_cuda_fwd_bwd(p), r32, r64, tag="cfg2/")


def test_config3_ou_tc_family_b4096():
    """BASELINE configs[2] through the tensor-core recurrence (VARIANT_TC, B = 4096 = 32 tiles, T = 100, C = 256):
    per-trajectory results on two 128-row slices (first tile, last tile) against fp64, and the complete iteration incl.
    every weight gradient on a 2-tile batch."""
    from viforsdes_b200 import _lib, ops

    ops.set_variant(_lib.VARIANT_TC)
    try:
        p = O.make_problem("ou", 4096, 100, **FULL)
        out = run_cuda_fwd_bwd(p)
        _check_slice(out, p, slice(0, 128), "cfg3/tc/b4096/rows0-127/")
        _check_slice(out, p, slice(4096 - 128, 4096), "cfg3/tc/b4096/rows3968-4095/")
        small = O.make_problem("ou", 256, 100, **FULL)
        r32, r64 = oracle_refs(small)
        check_iteration(run_cuda_fwd_bwd(small), r32, r64, tag="cfg3/tc/b256/")
    finally:
        ops.set_variant(_lib.VARIANT_AUTO)


def test_config4_lv_t20000_slice():
    """BASELINE configs[3] (dt = 0.002, T = 20 000, B = 128) on the GPU; one trajectory against the oracle (a stepwise
    fp64 + fp32 pass with autograd over 20 000 steps is ~1 min of CPU)."""
    p = O.make_problem("lv", 128, 20000, dt=0.002, **FULL)
    out = run_cuda_fwd_bwd(p)
    _check_slice(out, p, slice(5, 6), "cfg4/row5/")
    assert all(torch.isfinite(g).all() for g in out[4].values())


def test_config5_l96_b8192_user_sde():
    """BASELINE configs[4] at size through the entry point bench.py times (PathIteration, user SDE in PyTorch): rows of
    the first and last tile against the oracle, and a complete 384-trajectory iteration incl. every weight gradient."""
    from tests.test_gpu_config5 import _inputs_from_problem
    from viforsdes_b200.runner import PathIteration

    p = O.make_problem("l96", 8192, 100, **FULL)
    it = PathIteration(_inputs_from_problem(p), "cuda")
    it.step()
    res = it.results()
    out = (res["paths"], res["means"], res["chol"], res["terms"], res["grads"])
    _check_slice(out, p, slice(0, 64), "cfg5/b8192/rows0-63/")
    _check_slice(out, p, slice(8192 - 64, 8192), "cfg5/b8192/rows8128-8191/")
    del it
    small = O.make_problem("l96", 384, 100, **FULL)
    r32, r64 = oracle_refs(small)
    it = PathIteration(_inputs_from_problem(small), "cuda")
    it.step()
    res = it.results()
    check_iteration((res["paths"], res["means"], res["chol"], res["terms"], res["grads"]), r32, r64, tag="cfg5/b384/")


@pytest.mark.parametrize("kind,S,obs_dim", [("l96", 3, 2), ("lv", 2, 1), ("l96", 10, 4)])
def test_observation_matrix_branch(kind, S, obs_dim):
    """GaussianObservationLikelihood with obs_matrix [obs_dim, S] (core/observations.py:58-66): terms and every gradient."""
    from tests._util import build_head, cuda_inputs, cuda_sde, head_grads
    from viforsdes_b200.elbo import path_elbo_terms
    from viforsdes_b200.observations import GaussianObservationLikelihood, Observations
    from viforsdes_b200.state_space import StateSpace
    from viforsdes_b200.types import DiffusionPathSample

    p = O.make_problem(kind, 5, 40, context_dim=16, hidden_dim=32, num_layers=2, state_dim=S if kind == "l96" else None)
    g = torch.Generator().manual_seed(3)
    H = torch.randn(obs_dim, S, generator=g)
    values = torch.randn(p.obs_times.shape[0], obs_dim, generator=g) * 2 + (8.0 if kind == "l96" else 50.0)

    def oracle(dtype):
        c = lambda t: t.to(dtype)  # noqa: E731
        w = p.weights.map(lambda t: c(t).clone().requires_grad_(True))
        x0, ctx, th = (c(t).clone().requires_grad_(True) for t in (p.x0, p.context, p.theta))
        paths, means, chol = O.sample_paths(w, x0, ctx, th, c(p.eps), p.dt)
        t = O.elbo_terms(p.sde, paths, means, chol, th, p.dt, p.positive_dims, c(p.obs_times), c(values), p.obs_variance, c(H))
        (-t.path_elbo()).backward()
        grads = {"x0": x0.grad, "context": ctx.grad, "theta": th.grad, "out_w": w.out_w.grad, "w_ih_l0": w.w_ih[0].grad,
                 "w_hh_l1": w.w_hh[1].grad}
        return t, grads

    t32, g32 = oracle(torch.float32)
    t64, g64 = oracle(torch.float64)
    head = build_head(p)
    x0, full, view, theta, eps = cuda_inputs(p)
    paths, means, chol = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
    sample = DiffusionPathSample(paths, means, chol, StateSpace(S, list(p.positive_dims)))
    terms = path_elbo_terms(cuda_sde(p), Observations(times=p.obs_times, values=values),
                            GaussianObservationLikelihood(variance=p.obs_variance, obs_matrix=H), theta, sample, p.dt)
    (-(terms[:, 0] + terms[:, 1] - terms[:, 2] + terms[:, 3]).mean()).backward()
    assert_parity(terms[:, 0], t32.obs, t64.obs, name=f"obsmat/{kind}{S}/term_obs")
    assert_parity(terms[:, 1], t32.sde, t64.sde, name=f"obsmat/{kind}{S}/term_sde")
    hg = head_grads(head)
    got = {"x0": x0.grad, "context": full.grad[:, :40], "theta": theta.grad, "out_w": hg["out_w"], "w_ih_l0": hg["w_ih_l0"],
           "w_hh_l1": hg["w_hh_l1"]}
    for nm in g64:
        assert_parity(got[nm], g32[nm], g64[nm], name=f"obsmat/{kind}{S}/grad_{nm}")


def test_non_gaussian_likelihood_fallback():
    """A user ObservationLikelihood (protocol, core/observations.py:31-38) is evaluated in PyTorch on x[:, obs_idx] while
    the sde / gen / jacobian terms stay in the kernel (elbo.py): terms and gradients against the oracle with the same
    likelihood."""
    from tests._util import build_head, cuda_inputs, cuda_sde
    from viforsdes_b200.elbo import path_elbo_terms
    from viforsdes_b200.observations import Observations
    from viforsdes_b200.state_space import StateSpace
    from viforsdes_b200.types import DiffusionPathSample

    class Laplace:
        def log_prob(self, observations, state):
            return (-(observations - state).abs() / 0.7 - 0.3365).sum(dim=-1)

    p = O.make_problem("lv", 4, 30, context_dim=16, hidden_dim=32, num_layers=1)

    def oracle(dtype):
        c = lambda t: t.to(dtype)  # noqa: E731
        w = p.weights.map(lambda t: c(t).clone().requires_grad_(True))
        x0, ctx, th = (c(t).clone().requires_grad_(True) for t in (p.x0, p.context, p.theta))
        paths, means, chol = O.sample_paths(w, x0, ctx, th, c(p.eps), p.dt)
        t = O.elbo_terms(p.sde, paths, means, chol, th, p.dt, p.positive_dims, c(p.obs_times), c(p.obs_values), p.obs_variance)
        idx = O.obs_indices(p.obs_times, p.dt, 30)
        x = O.to_state(paths, p.positive_dims)
        obs = Laplace().log_prob(c(p.obs_values)[None].expand(4, -1, -1), x[:, idx]).sum(-1)
        (-(obs + t.sde - t.gen + t.jac).mean()).backward()
        return obs, {"x0": x0.grad, "context": ctx.grad, "theta": th.grad}

    o32, g32 = oracle(torch.float32)
    o64, g64 = oracle(torch.float64)
    head = build_head(p)
    x0, full, view, theta, eps = cuda_inputs(p)
    paths, means, chol = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
    sample = DiffusionPathSample(paths, means, chol, StateSpace(2, [0, 1]))
    terms = path_elbo_terms(cuda_sde(p), Observations(times=p.obs_times, values=p.obs_values), Laplace(), theta, sample, p.dt)
    (-(terms[:, 0] + terms[:, 1] - terms[:, 2] + terms[:, 3]).mean()).backward()
    assert_parity(terms[:, 0], o32, o64, name="laplace/term_obs")
    for nm, a in (("x0", x0.grad), ("context", full.grad[:, :30]), ("theta", theta.grad)):
        assert_parity(a, g32[nm], g64[nm], name=f"laplace/grad_{nm}")


@pytest.mark.parametrize("E,Cd,B,T", [(24, 16, 5, 13), (128, 128, 3, 40)])
def test_context_fold_matches_oracle(E, Cd, B, T):
    """Context-producer fold against the ORACLE: the reference computes context = tokens . W_op^T + b_op
    (primitives/sit.py:156-158) and feeds it to the head; here the fold is evaluated in fp64 / fp32 on the CPU and the
    folded CUDA call must reproduce paths and the gradients of both factors."""
    from torch import nn

    from tests._util import build_head

    p = O.make_problem("lv", B, T, context_dim=Cd, hidden_dim=64, num_layers=2)
    g = torch.Generator().manual_seed(E)
    tokens = 0.3 * torch.randn(B, T, E, generator=g)
    proj = nn.Linear(E, Cd)
    cts = [torch.randn(B, T + 1, 2, generator=g), torch.randn(B, T, 2, generator=g), torch.randn(B, T, 2, 2, generator=g)]

    def oracle(dtype):
        c = lambda t: t.detach().to(dtype).clone().requires_grad_(True)  # noqa: E731
        w = p.weights.map(c)
        tk, W, b, th = c(tokens), c(proj.weight), c(proj.bias), c(p.theta)
        out = O.sample_paths(w, p.x0.to(dtype), tk @ W.T + b, th, p.eps.to(dtype), p.dt)
        sum((o * ct.to(dtype)).sum() for o, ct in zip(out, cts)).backward()
        return out, {"tokens": tk.grad, "proj_w": W.grad, "proj_b": b.grad, "theta": th.grad, "w_ih_l0": w.w_ih[0].grad,
                     "b_ih_l0": w.b_ih[0].grad, "w_hh_l1": w.w_hh[1].grad, "out_w": w.out_w.grad}

    o32, g32 = oracle(torch.float32)
    o64, g64 = oracle(torch.float64)
    head = build_head(p)
    pr = nn.Linear(E, Cd).cuda()
    pr.load_state_dict(proj.state_dict())
    tk = tokens.cuda().requires_grad_(True)
    th = p.theta.cuda().requires_grad_(True)
    out = head.sample_diffusion_paths_from_tokens(p.x0.cuda(), tk, pr, th, p.eps.cuda(), p.dt)
    sum((o * ct.cuda()).sum() for o, ct in zip(out, cts)).backward()
    for a, b32, b64, nm in zip(out, o32, o64, ("paths", "means", "chol")):
        assert_parity(a, b32, b64, name=f"fold/E{E}/{nm}")
    got = {"tokens": tk.grad, "proj_w": pr.weight.grad, "proj_b": pr.bias.grad, "theta": th.grad,
           "w_ih_l0": head.gru.weight_ih_l0.grad, "b_ih_l0": head.gru.bias_ih_l0.grad,
           "w_hh_l1": head.gru.weight_hh_l1.grad, "out_w": head.out_proj.weight.grad}
    for nm in g64:
        assert_parity(got[nm], g32[nm], g64[nm], name=f"fold/E{E}/grad_{nm}")
