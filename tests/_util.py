"""Shared helpers for the GPU parity tests: run the CUDA path on an oracle Problem."""
from __future__ import annotations

import torch

from oracle import oracle_torch as O


def normwise(a: torch.Tensor, ref: torch.Tensor) -> float:
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    den = ref.abs().max().item()
    return (a - ref).abs().max().item() / (den if den > 0 else 1.0)


def assert_close(a, ref, rtol=1e-4, atol_scale=1e-5, name=""):
    """|a - ref| <= atol_scale * max|ref| + rtol * |ref| elementwise (fp32 parity bar of
    BASELINE.json: rtol 1e-4, with an absolute floor scaled to the tensor's magnitude)."""
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    assert a.shape == ref.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(ref.shape)}"
    scale = ref.abs().max().item() if ref.numel() else 0.0
    err = (a - ref).abs()
    tol = atol_scale * scale + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), (f"{name}: {int(bad.sum())}/{a.numel()} elements off; max abs err "
                           f"{err.max().item():.3e} (scale {scale:.3e}, normwise {err.max().item() / (scale or 1):.3e})")


# achieved errors of every assert_parity call: tests/conftest.py writes them to gpurun_out/parity_log.jsonl at session end,
# tools/parity_table.py turns that into profiles/parity_r2.md
PARITY_LOG: list = []
NOISE_CAP = 1e-3  # the fp32-noise widening never exceeds this fraction of the tensor's magnitude


def assert_parity(a, ref32, ref64, rtol=1e-4, atol_scale=1e-5, noise_mult=3.0, name="", noise_cap=NOISE_CAP, atol_abs=0.0):
    """Parity against the fp64 oracle with the fp32 bar of BASELINE.json (rtol 1e-4 elementwise plus an
    absolute floor of atol_scale x max|ref|), widened by `noise_mult` x the rounding noise the
    reference's own fp32 path shows on this tensor (max|ref32 - ref64|), the widening capped at
    `noise_cap` x max|ref|.  The widening matters for Lotka-Volterra-like problems (|z| ~ 70-450, saturated
    gates, A_ii = L_ii sqrt(dt) ~ 1e-3): there the ELBO cotangents are differences of O(1/A_ii^2) terms and
    any two fp32 implementations -- including the reference's PyTorch path vs its own Triton kernels --
    differ by more than 1e-4 relative.  The achieved normwise error is recorded in PARITY_LOG."""
    a, r32, r64 = (t.detach().double().cpu() for t in (a, ref32, ref64))
    assert a.shape == r64.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(r64.shape)}"
    if a.numel() == 0:
        return
    scale = r64.abs().max().item()
    noise = (r32 - r64).abs().max().item()
    err = (a - r64).abs()
    widen = min(noise_mult * noise, noise_cap * scale)
    tol = rtol * r64.abs() + atol_scale * scale + widen + atol_abs  # atol_abs: floor for terms that are physically zero
    PARITY_LOG.append({"name": name, "numel": a.numel(), "scale": scale, "max_abs_err": err.max().item(),
                       "normwise_err": err.max().item() / (scale or 1.0), "fp32_ref_noise_normwise": noise / (scale or 1.0),
                       "widening_normwise": widen / (scale or 1.0)})
    bad = err > tol
    assert not bad.any(), (f"{name}: {int(bad.sum())}/{a.numel()} elements off; max abs err {err.max().item():.3e}, "
                           f"scale {scale:.3e}, fp32-reference noise {noise:.3e}")


def build_head(p: O.Problem, device="cuda"):
    from viforsdes_b200.head import DiffusionTransitionHead, HeadConfig

    w = p.weights
    head = DiffusionTransitionHead(w.state_dim, w.context_dim, w.param_dim,
                                   HeadConfig(hidden_dim=w.hidden_dim, num_layers=w.num_layers))
    with torch.no_grad():
        for k in range(w.num_layers):
            getattr(head.gru, f"weight_ih_l{k}").copy_(w.w_ih[k])
            getattr(head.gru, f"weight_hh_l{k}").copy_(w.w_hh[k])
            getattr(head.gru, f"bias_ih_l{k}").copy_(w.b_ih[k])
            getattr(head.gru, f"bias_hh_l{k}").copy_(w.b_hh[k])
        head.out_proj.weight.copy_(w.out_w)
        head.out_proj.bias.copy_(w.out_b)
    return head.to(device).train()


def cuda_sde(p: O.Problem):
    from viforsdes_b200 import sde as vs

    if p.name == "ou":
        return vs.OrnsteinUhlenbeck()
    if p.name == "lv":
        return vs.LotkaVolterra()
    return p.sde  # user-defined SDE through the protocol (PyTorch drift / diffusion)


def head_grads(head) -> dict:
    g = {}
    for k in range(head.num_layers):
        g[f"w_ih_l{k}"] = getattr(head.gru, f"weight_ih_l{k}").grad
        g[f"w_hh_l{k}"] = getattr(head.gru, f"weight_hh_l{k}").grad
        g[f"b_ih_l{k}"] = getattr(head.gru, f"bias_ih_l{k}").grad
        g[f"b_hh_l{k}"] = getattr(head.gru, f"bias_hh_l{k}").grad
    g["out_w"] = head.out_proj.weight.grad
    g["out_b"] = head.out_proj.bias.grad
    return g


def cuda_inputs(p: O.Problem, ctx_dtype=torch.float32, strided=True):
    """x0, context (strided [B,T,C] view of a [B,T+1,C] leaf, like the reference call site), theta, eps."""
    dev = "cuda"
    B, T, Cd = p.context.shape
    x0 = p.x0.to(dev).requires_grad_(True)
    theta = p.theta.to(dev).requires_grad_(True)
    eps = p.eps.to(dev)
    if strided:
        full = torch.zeros(B, T + 1, Cd, device=dev, dtype=ctx_dtype)
        full[:, :-1] = p.context.to(dev).to(ctx_dtype)
        full.requires_grad_(True)
        view = full[:, :-1]
    else:
        full = p.context.to(dev).to(ctx_dtype).contiguous().requires_grad_(True)
        view = full
    return x0, full, view, theta, eps


def run_cuda_fwd_bwd(p: O.Problem, ctx_dtype=torch.float32, strided=True):
    """CUDA iteration through the public modules: head.sample_diffusion_paths -> path_elbo_terms ->
    backward of -mean(obs + sde - gen + jac).  Mirrors oracle.run_fwd_bwd."""
    from viforsdes_b200.elbo import path_elbo_terms
    from viforsdes_b200.observations import GaussianObservationLikelihood, Observations
    from viforsdes_b200.state_space import StateSpace
    from viforsdes_b200.types import DiffusionPathSample

    head = build_head(p)
    x0, full, view, theta, eps = cuda_inputs(p, ctx_dtype, strided)
    paths, means, chol = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
    sample = DiffusionPathSample(paths, means, chol, StateSpace(p.weights.state_dim, list(p.positive_dims)))
    terms = path_elbo_terms(cuda_sde(p), Observations(times=p.obs_times, values=p.obs_values),
                            GaussianObservationLikelihood(variance=p.obs_variance), theta, sample, p.dt)
    loss = -(terms[:, 0] + terms[:, 1] - terms[:, 2] + terms[:, 3]).mean()
    loss.backward()
    B, T, _ = p.context.shape
    grads = {"x0": x0.grad, "context": full.grad[:, :T] if strided else full.grad, "theta": theta.grad}
    grads.update(head_grads(head))
    return paths.detach(), means.detach(), chol.detach(), terms.detach(), grads


def oracle_refs(p: O.Problem):
    """(fp32, fp64) oracle iterations on the same inputs."""
    return O.run_fwd_bwd(p), O.run_fwd_bwd(p, dtype=torch.float64)


def check_iteration(cuda_out, r32, r64, tag="", noise_cap=NOISE_CAP):
    """paths / means / chol / 4 ELBO terms / every gradient against the oracle pair."""
    paths, means, chol, terms, grads = cuda_out
    for a, b32, b64, nm in zip((paths, means, chol), r32[:3], r64[:3], ("paths", "means", "chol")):
        assert_parity(a, b32, b64, name=f"{tag}{nm}", noise_cap=noise_cap)
    tscale = max(getattr(r64[3], nm).abs().max().item() for nm in ("obs", "sde", "gen", "jac"))
    for j, nm in enumerate(("obs", "sde", "gen", "jac")):
        # the four terms are summed into one ELBO: a term that is physically zero (log-Jacobian of a saturated softplus,
        # ~1e-24) is held to the magnitude of the sum, not to its own
        assert_parity(terms[:, j], getattr(r32[3], nm), getattr(r64[3], nm), name=f"{tag}term_{nm}", noise_cap=noise_cap,
                      atol_abs=1e-7 * tscale)
    for nm in r64[4]:
        assert_parity(grads[nm], r32[4][nm], r64[4][nm], name=f"{tag}grad_{nm}", noise_cap=noise_cap)
