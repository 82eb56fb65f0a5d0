"""BASELINE.json configs[4] (user-defined SDE: 10-D stochastic Lorenz-96) through the two entry points bench.py
times -- the device-resident PathIteration (`value` leg) and the host-buffer C-ABI session with its
visde_user_sde hooks (`e2e` leg) -- against the CPU oracle on the same seeded inputs."""
from __future__ import annotations

import pytest
import torch

from oracle import oracle_torch as O
from tests._util import assert_parity, oracle_refs

pytestmark = pytest.mark.gpu


def _inputs_from_problem(p: O.Problem):
    """synthetic.Inputs carrying the oracle problem's tensors (what bench.py builds with make_inputs)."""
    from viforsdes_b200 import _lib
    from viforsdes_b200.synthetic import Inputs, Lorenz96

    B, T, Cd = p.context.shape
    full = torch.zeros(B, T + 1, Cd)
    full[:, :T] = p.context
    w = p.weights
    kind = {"ou": _lib.SDE_OU, "lv": _lib.SDE_LV}.get(p.name, _lib.SDE_GENERIC)
    sde = Lorenz96(w.state_dim) if p.name == "l96" else None
    return Inputs(p.name, kind, tuple(p.positive_dims), p.x0, full, p.theta, p.eps, w.w_ih, w.w_hh, w.b_ih, w.b_hh,
                  w.out_w, w.out_b, p.dt, p.obs_times, p.obs_values, p.obs_variance, sde)


def _check(res, r32, r64, tag):
    for j, nm in enumerate(("obs", "sde", "gen", "jac")):
        assert_parity(res["terms"][:, j], getattr(r32[3], nm), getattr(r64[3], nm), name=f"{tag}term_{nm}")
    for nm in r64[4]:
        if nm in res["grads"]:
            assert_parity(res["grads"][nm], r32[4][nm], r64[4][nm], name=f"{tag}grad_{nm}")


@pytest.mark.parametrize("B,T,S,Cd", [(5, 21, 10, 32), (130, 12, 10, 128), (3, 9, 4, 16)])
def test_path_iteration_user_sde_matches_oracle(B, T, S, Cd):
    from viforsdes_b200.runner import PathIteration

    p = O.make_problem("l96", B, T, context_dim=Cd, hidden_dim=64, num_layers=2, state_dim=S)
    r32, r64 = oracle_refs(p)
    it = PathIteration(_inputs_from_problem(p), "cuda")
    it.step()
    res = it.results()
    for a, b32, b64, nm in zip((res["paths"], res["means"], res["chol"]), r32[:3], r64[:3], ("paths", "means", "chol")):
        assert_parity(a, b32, b64, name=nm)
    _check(res, r32, r64, "eager/")
    # the captured graph (PyTorch drift / diffusion + their autograd VJP inside it) replays to the same numbers
    first = {k: v.clone() for k, v in res["grads"].items()}
    it.capture()
    it.replay()
    torch.cuda.synchronize()
    res2 = it.results()
    for k, v in first.items():
        assert torch.equal(res2["grads"][k], v), f"graph replay differs from the eager iteration ({k})"


def test_path_iteration_user_sde_with_positive_dims():
    """A user SDE on a softplus-transformed state: the chain g_drift -> g_x -> g_z goes through d softplus / dz."""
    from viforsdes_b200 import _lib
    from viforsdes_b200.runner import PathIteration

    p = O.make_problem("lv", 4, 15, context_dim=16, hidden_dim=64, num_layers=2)
    r32, r64 = oracle_refs(p)
    inp = _inputs_from_problem(p)
    inp.sde_kind, inp.sde = _lib.SDE_GENERIC, O.LotkaVolterra()  # the LV model through the user-SDE route
    it = PathIteration(inp, "cuda")
    it.step()
    _check(it.results(), r32, r64, "lv-as-user-sde/")


@pytest.mark.parametrize("ctx_dtype", [torch.float32, torch.bfloat16])
def test_host_session_user_sde_matches_oracle(ctx_dtype):
    from viforsdes_b200.session import HostSession

    p = O.make_problem("l96", 6, 17, context_dim=32, hidden_dim=64, num_layers=2, state_dim=10)
    if ctx_dtype == torch.bfloat16:  # the oracle sees the bf16-rounded context the session is handed
        p.context = p.context.to(torch.bfloat16).to(torch.float32)
    r32, r64 = oracle_refs(p)
    sess = HostSession.from_problem(p, want_grad_context=True, context_dtype=ctx_dtype)
    res = sess.step()
    if ctx_dtype == torch.bfloat16:
        g = res["grads"].pop("context")
        assert g.dtype == torch.bfloat16
        ref = r64[4]["context"]
        assert (g.double() - ref).abs().max() <= 2.0 ** -8 * ref.abs().max() + 1e-12, "bf16 grad_context beyond bf16 rounding"
    _check(res, r32, r64, f"session/{ctx_dtype}/")
    # pipelined form: same numbers, two in flight
    ref = {k: v.clone() for k, v in res["grads"].items()}
    sess.submit()
    sess.submit()
    for _ in range(2):
        out = sess.wait()
        for k, v in ref.items():
            assert torch.equal(out["grads"][k], v), k
    assert sess.h2d_bytes > 0 and sess.launches > 0
    # from the 4th call on the user-SDE hooks replay as CUDA graphs: same numbers
    for _ in range(3):
        out = sess.step()
    assert sess._hooks.g_eval is not None and sess._hooks.g_vjp is not None, "hooks should have been captured"
    for k, v in ref.items():
        assert torch.equal(out["grads"][k], v), f"graph-replayed hooks changed {k}"
    # changed inputs flow through the captured hooks
    sess.theta.mul_(1.01)
    moved = sess.step()
    assert not torch.equal(moved["grads"]["theta"], ref["theta"])
    sess.close()


def test_host_session_user_sde_hook_errors_propagate():
    from viforsdes_b200.session import HostSession

    class Broken(O.Lorenz96):
        def diffusion(self, x, p):
            raise FloatingPointError("user diffusion failed")

    p = O.make_problem("l96", 2, 5, context_dim=8, hidden_dim=64, num_layers=1, state_dim=5)
    p.sde = Broken(5)
    sess = HostSession.from_problem(p)
    with pytest.raises(FloatingPointError):
        sess.step()
    sess.close()
    with pytest.raises(ValueError):  # GENERIC without hooks is refused by the C ABI wrapper
        HostSession.from_problem(O.make_problem("l96", 2, 5, context_dim=8, hidden_dim=64, num_layers=1, state_dim=5), sde=None)


@pytest.mark.parametrize("name,S", [("l96", 10), ("ou", 1)])
def test_host_session_device_noise_matches_host_fed_philox(name, S):
    """eps == NULL: iteration i draws visde_philox_normal(seed + i) on the device (diffusion_path_sampler.py:57 draws on the
    device too); the same draws fed from the host give bit-identical outputs, and eps leaves the H2D byte count."""
    from viforsdes_b200.euler_maruyama import philox_normal
    from viforsdes_b200.session import HostSession

    B, T, seed = 6, 17, 1234
    kw = dict(context_dim=32, hidden_dim=64, num_layers=2)
    p = O.make_problem(name, B, T, state_dim=S, **kw) if name == "l96" else O.make_problem(name, B, T, **kw)
    dev = HostSession.from_problem(p, device_noise_seed=seed)
    outs = []
    for _ in range(3):  # ou: the third step replays the captured graph behind a fresh draw
        o = dev.step()  # views of the session's two host output sets: copy before they are reused
        outs.append({k: v.clone() for k, v in o["grads"].items()} | {"terms": o["terms"].clone()})
    fed = HostSession.from_problem(p)
    assert fed.h2d_bytes - dev.h2d_bytes == 4 * B * T * S
    for i, o in enumerate(outs):
        fed.eps.copy_(philox_normal(seed + i, B, T, S).cpu())
        r = fed.step()
        assert torch.equal(r["terms"], o["terms"]), f"iteration {i}: terms differ"
        for k, v in r["grads"].items():
            assert torch.equal(v, o[k]), f"iteration {i}: {k} differs"
    assert not torch.equal(outs[0]["terms"], outs[1]["terms"]), "every iteration must draw fresh noise"
    # and the host-fed run agrees with the oracle on those draws
    p.eps = philox_normal(seed + 2, B, T, S).cpu()
    r32, r64 = oracle_refs(p)
    _check(r, r32, r64, "device-noise/")
    dev.close()
    fed.close()
