"""Generate golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference/src`` (read-only, absent on the GPU box), runs
  (A) ``DiffusionTransitionHead.forward`` stepwise (models/head.py:68-86) +
      ``compute_evidence_lower_bound`` (inference/evidence_lower_bound.py:19-74) with
      autograd gradients, and
  (B) the reference's own Triton kernels ``launch_fwd``/``launch_bwd``
      (kernels/forward.py:378, kernels/backward.py:627) on CPU under
      ``TRITON_INTERPRET=1`` with an external tanh shim,
on the seeded inputs of ``oracle.oracle_torch.make_problem`` and writes small ``*.pt``
fixtures next to this file.  The fixtures pin the oracle (tests/test_oracle_golden.py)
and, on the GPU, the CUDA path (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path

os.environ.setdefault("TRITON_INTERPRET", "1")

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "src"))

# matplotlib is not installed here; the reference only needs the names at import time
for name, attrs in {
    "matplotlib": {},
    "matplotlib.pyplot": {},
    "matplotlib.axes": {"Axes": object},
    "matplotlib.figure": {"Figure": object},
}.items():
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)

import torch  # noqa: E402

from oracle import oracle_torch as O  # noqa: E402

from variational_sde.config import HeadConfig  # noqa: E402
from variational_sde.core.observations import GaussianObservationLikelihood, Observations  # noqa: E402
from variational_sde.core.priors import Prior, PriorType  # noqa: E402
from variational_sde.inference.evidence_lower_bound import compute_evidence_lower_bound  # noqa: E402
from variational_sde.inference.state_space import StateSpace  # noqa: E402
from variational_sde.inference.types import DiffusionPathSample  # noqa: E402
from variational_sde.models.head import DiffusionTransitionHead  # noqa: E402
from variational_sde.models.sde_parameter_posterior import SDEParameterPosterior  # noqa: E402


def _load_example(name: str):
    spec = importlib.util.spec_from_file_location(name, REF / "examples" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_head(p: O.Problem) -> DiffusionTransitionHead:
    w = p.weights
    head = DiffusionTransitionHead(
        w.state_dim, w.context_dim, w.param_dim, HeadConfig(hidden_dim=w.hidden_dim, num_layers=w.num_layers)
    )
    with torch.no_grad():
        for k in range(w.num_layers):
            getattr(head.gru, f"weight_ih_l{k}").copy_(w.w_ih[k])
            getattr(head.gru, f"weight_hh_l{k}").copy_(w.w_hh[k])
            getattr(head.gru, f"bias_ih_l{k}").copy_(w.b_ih[k])
            getattr(head.gru, f"bias_hh_l{k}").copy_(w.b_hh[k])
        head.out_proj.weight.copy_(w.out_w)
        head.out_proj.bias.copy_(w.out_b)
    return head


def ref_sde(kind: str, state_dim: int):
    if kind == "ou":
        return _load_example("ornstein_uhlenbeck").OrnsteinUhlenbeck()
    if kind == "lv":
        return _load_example("lotka_volterra").LotkaVolterra()
    return O.Lorenz96(state_dim)  # user-defined SDE through the reference's SDE protocol


def golden_stepwise(kind: str, B: int, T: int, **kw) -> dict:
    p = O.make_problem(kind, B, T, **kw)
    head = ref_head(p)
    S = p.weights.state_dim
    x0 = p.x0.clone().requires_grad_(True)
    ctx = p.context.clone().requires_grad_(True)
    theta = p.theta.clone().requires_grad_(True)
    hidden = head.init_hidden(B, torch.device("cpu"))
    z = x0
    paths, means, chols = [z], [], []
    for t in range(T):
        mu, L, hidden = head(z, ctx[:, t], theta, hidden)
        z = z + mu * p.dt + torch.einsum("bij,bj->bi", L, p.eps[:, t]) * (p.dt**0.5)
        paths.append(z)
        means.append(mu)
        chols.append(L)
    paths_t, means_t, chol_t = torch.stack(paths, 1), torch.stack(means, 1), torch.stack(chols, 1)
    sample = DiffusionPathSample(
        z=paths_t, transition_means=means_t, transition_cholesky=chol_t,
        state_space=StateSpace(S, list(p.positive_dims)),
    )
    P = p.weights.param_dim
    post = SDEParameterPosterior(P, [], init_mean=torch.zeros(P), init_std=2.0)
    prior = Prior(type=PriorType.NORMAL, mean=0.0, std=3.0, dim=P)
    res = compute_evidence_lower_bound(
        ref_sde(kind, S),
        Observations(times=p.obs_times, values=p.obs_values),
        GaussianObservationLikelihood(variance=p.obs_variance),
        prior, post, theta, sample, p.dt,
    )
    c = res.components
    path_elbo = res.evidence_lower_bound - c.prior_log_prob + c.posterior_log_prob
    jac_mean = path_elbo - c.observation_log_prob - c.sde_log_prob + c.generative_log_prob
    (-path_elbo).backward()
    out = {
        "kind": kind, "B": B, "T": T, "kw": kw,
        "paths": paths_t.detach(), "means": means_t.detach(), "chol": chol_t.detach(),
        "obs_mean": c.observation_log_prob.detach(), "sde_mean": c.sde_log_prob.detach(),
        "gen_mean": c.generative_log_prob.detach(), "jac_mean": jac_mean.detach(),
        "path_elbo": path_elbo.detach(),
        "g_x0": x0.grad, "g_context": ctx.grad, "g_theta": theta.grad,
        "g_out_w": head.out_proj.weight.grad, "g_out_b": head.out_proj.bias.grad,
    }
    for k in range(p.weights.num_layers):
        for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            out[f"g_{nm}_l{k}"] = getattr(head.gru, f"{nm}_l{k}").grad
    return out


def golden_triton(kind: str, B: int, T: int, **kw) -> dict:
    """The reference kernels themselves (interpreter mode) with N(0,1) upstream cotangents."""
    import triton.language as tl
    import variational_sde.kernels.backward as kb
    import variational_sde.kernels.forward as kf
    from variational_sde.kernels.weights import SDEWeights

    class _Shim:
        @staticmethod
        def tanh(x):
            return 2.0 * tl.sigmoid(2.0 * x) - 1.0

    kf.libdevice = _Shim
    if hasattr(kb, "libdevice"):
        kb.libdevice = _Shim

    p = O.make_problem(kind, B, T, **kw)
    head = ref_head(p)
    w = SDEWeights.from_modules(head.gru, head.out_proj, p.weights.context_dim, p.weights.param_dim, p.weights.state_dim)
    ctx = p.context.contiguous()
    paths, means, chol, saved = kf.launch_fwd(p.x0, ctx, p.theta, p.eps, w, p.dt, save_activations=True)
    g = torch.Generator().manual_seed(99)
    gP, gM, gL = torch.randn(paths.shape, generator=g), torch.randn(means.shape, generator=g), torch.randn(chol.shape, generator=g)
    grads = kb.launch_bwd(gP, gM, gL, ctx, p.theta, p.eps, saved, w, p.dt)
    names = ["g_x0", "g_context", "g_theta", "g_w_ih_l0", "g_w_hh_l0", "g_b_ih_l0", "g_b_hh_l0",
             "g_w_ih_stack", "g_w_hh_stack", "g_b_ih_stack", "g_b_hh_stack", "g_out_w", "g_out_b"]
    out = {"kind": kind, "B": B, "T": T, "kw": kw, "paths": paths, "means": means, "chol": chol,
           "gP": gP, "gM": gM, "gL": gL}
    out.update({n: t.clone() for n, t in zip(names, grads)})
    return out


CASES = {
    "stepwise_ou": ("ou", 4, 20, dict(dt=0.05, context_dim=16, hidden_dim=32, num_layers=2)),
    "stepwise_lv": ("lv", 3, 24, dict(dt=0.05, context_dim=8, hidden_dim=16, num_layers=1)),
    "stepwise_l96": ("l96", 2, 10, dict(dt=0.05, context_dim=8, hidden_dim=24, num_layers=3, state_dim=4)),
    "stepwise_ou_h64": ("ou", 2, 40, dict(dt=0.05, context_dim=32, hidden_dim=64, num_layers=2)),
}
TRITON_CASES = {
    "triton_lv": ("lv", 2, 5, dict(dt=0.05, context_dim=8, hidden_dim=16, num_layers=2)),
    "triton_l96": ("l96", 2, 4, dict(dt=0.05, context_dim=4, hidden_dim=16, num_layers=3, state_dim=3)),
}


def main() -> None:
    torch.manual_seed(0)
    for name, (kind, B, T, kw) in CASES.items():
        torch.save(golden_stepwise(kind, B, T, **kw), HERE / f"{name}.pt")
        print("wrote", name)
    for name, (kind, B, T, kw) in TRITON_CASES.items():
        torch.save(golden_triton(kind, B, T, **kw), HERE / f"{name}.pt")
        print("wrote", name)


if __name__ == "__main__":
    main()
