"""Direct parity against the UNMODIFIED reference running its own Triton kernels on the same GPU (when a copy of
the reference sources travelled with the snapshot as git-ignored baseline/_ref/src; skipped otherwise).  The CPU
oracle is pinned to reference outputs generated offline (tests/golden); this closes the loop on the target GPU:
same inputs, same weights, reference `DiffusionTransitionHead.sample_diffusion_paths` (models/head.py:156-209 ->
kernels/forward.py, kernels/backward.py) vs this library, forward outputs and every gradient."""
import sys
import types
from pathlib import Path

import pytest
import torch

from oracle import oracle_torch as O
from tests._util import build_head, cuda_inputs, head_grads, normwise

pytestmark = pytest.mark.gpu
REF_SRC = Path(__file__).resolve().parents[1] / "baseline" / "_ref" / "src"


def _reference_head(p):
    if not REF_SRC.exists():
        pytest.skip("baseline/_ref/src (copy of the reference sources) is not present")
    if str(REF_SRC) not in sys.path:
        sys.path.insert(0, str(REF_SRC))
    for name, attrs in {"matplotlib": {}, "matplotlib.pyplot": {}, "matplotlib.axes": {"Axes": object},
                        "matplotlib.figure": {"Figure": object}}.items():
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
    try:
        from variational_sde.config import HeadConfig
        from variational_sde.models.head import DiffusionTransitionHead as RefHead
    except Exception as e:  # pragma: no cover - environment without the reference's dependencies
        pytest.skip(f"reference not importable here: {e}")
    w = p.weights
    ref = RefHead(w.state_dim, w.context_dim, w.param_dim, HeadConfig(hidden_dim=w.hidden_dim, num_layers=w.num_layers))
    with torch.no_grad():
        for k in range(w.num_layers):
            getattr(ref.gru, f"weight_ih_l{k}").copy_(w.w_ih[k])
            getattr(ref.gru, f"weight_hh_l{k}").copy_(w.w_hh[k])
            getattr(ref.gru, f"bias_ih_l{k}").copy_(w.b_ih[k])
            getattr(ref.gru, f"bias_hh_l{k}").copy_(w.b_hh[k])
        ref.out_proj.weight.copy_(w.out_w)
        ref.out_proj.bias.copy_(w.out_b)
    return ref.cuda().train()


def _run(head, p, cts):
    x0, full, view, theta, eps = cuda_inputs(p)
    out = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
    torch.autograd.backward(list(out), cts if cts is not None else [torch.ones_like(o) for o in out])
    torch.cuda.synchronize()
    return [o.detach() for o in out], {"x0": x0.grad, "context": full.grad, "theta": theta.grad, **head_grads(head)}


# both cases share the reference's Triton specialisation (S, C, H, NL are constexpr there): one ~1 min JIT
@pytest.mark.parametrize("kind,B,T,kw", [
    ("lv", 6, 33, dict(context_dim=128, hidden_dim=64, num_layers=2)),      # register-resident family + tcgen05 GEMMs
    ("lv", 130, 12, dict(context_dim=128, hidden_dim=64, num_layers=2)),    # + the tensor-core recurrence family
])
def test_matches_reference_triton_kernels(kind, B, T, kw):
    from viforsdes_b200 import _lib, ops

    p = O.make_problem(kind, B, T, **kw)
    ref = _reference_head(p)
    g = torch.Generator(device="cuda").manual_seed(5)
    S = p.weights.state_dim
    cts = [torch.randn(B, T + 1, S, device="cuda", generator=g), torch.randn(B, T, S, device="cuda", generator=g),
           torch.randn(B, T, S, S, device="cuda", generator=g)]
    try:
        r_out, r_grads = _run(ref, p, cts)
    except Exception as e:  # pragma: no cover - Triton cannot JIT here
        pytest.skip(f"reference Triton kernels did not run here: {e}")
    for v in ([_lib.VARIANT_AUTO, _lib.VARIANT_TC] if B >= 128 else [_lib.VARIANT_AUTO]):
        ops.set_variant(v)
        o_out, o_grads = _run(build_head(p), p, cts)
        ops.set_variant(_lib.VARIANT_AUTO)
        for a, b, nm in zip(o_out, r_out, ("paths", "means", "chol")):
            assert normwise(a, b) < 1e-4, f"v{v} {nm}: {normwise(a, b)}"
        # the reference accumulates its weight gradients with fp32 atomics in arbitrary order: compare normwise
        for k in o_grads:
            assert normwise(o_grads[k], r_grads[k]) < 2e-4, f"v{v} grad_{k}: {normwise(o_grads[k], r_grads[k])}"
