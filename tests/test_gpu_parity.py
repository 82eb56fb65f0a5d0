"""GPU parity tests: the CUDA path (through the torch.library ops -> ctypes -> C ABI) against the
CPU oracle on the same seeded inputs with injected noise, and against the committed reference
outputs in tests/golden/.  Bar (BASELINE.json): rtol 1e-4 in FP32 (elementwise, with an absolute
floor of 1e-5 x the tensor's max magnitude)."""
from __future__ import annotations

from pathlib import Path

import pytest
import torch

from oracle import oracle_torch as O
from tests._util import (assert_close, assert_parity, build_head, check_iteration, cuda_inputs, head_grads, normwise,
                         oracle_refs, run_cuda_fwd_bwd)

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"

# (kind, B, T, kwargs): covers both kernel families, padding (H=48, 20), NL 1..4, S 1..10
CASES = {
    "ou_h32_l2": ("ou", 5, 23, dict(context_dim=16, hidden_dim=32, num_layers=2)),
    "lv_h16_l1": ("lv", 3, 24, dict(context_dim=8, hidden_dim=16, num_layers=1)),
    "ou_h64_l2": ("ou", 4, 40, dict(context_dim=32, hidden_dim=64, num_layers=2)),
    "lv_h64_l2": ("lv", 6, 33, dict(context_dim=64, hidden_dim=64, num_layers=2)),
    "lv_h48_l2": ("lv", 3, 17, dict(context_dim=24, hidden_dim=48, num_layers=2)),
    "ou_h20_l1": ("ou", 2, 9, dict(context_dim=5, hidden_dim=20, num_layers=1)),
    "l96s3_h64_l2": ("l96", 3, 12, dict(context_dim=16, hidden_dim=64, num_layers=2, state_dim=3)),
    "l96s4_h24_l3": ("l96", 2, 10, dict(context_dim=8, hidden_dim=24, num_layers=3, state_dim=4)),
    "l96s10_h64_l2": ("l96", 3, 11, dict(context_dim=32, hidden_dim=64, num_layers=2, state_dim=10)),
    "l96s5_h96_l4": ("l96", 2, 7, dict(context_dim=12, hidden_dim=96, num_layers=4, state_dim=5)),
    "ou_h130_l2": ("ou", 2, 6, dict(context_dim=7, hidden_dim=130, num_layers=2)),
    # tensor-core GEMM eligible shapes (H = 64, C in {128, 256}); T > 128 spans several row tiles
    "lv_h64_c128_l2": ("lv", 3, 37, dict(context_dim=128, hidden_dim=64, num_layers=2)),
    "ou_h64_c256_l1": ("ou", 2, 150, dict(context_dim=256, hidden_dim=64, num_layers=1)),
    "lv_h64_c256_l2": ("lv", 2, 300, dict(context_dim=256, hidden_dim=64, num_layers=2)),
    "l96s6_h64_c128_l2": ("l96", 2, 40, dict(context_dim=128, hidden_dim=64, num_layers=2, state_dim=6)),
}
# tensor-core recurrence eligible (H = 64, NL <= 2, S <= 4, tcgen05 K0); two 128-trajectory tiles, the second ragged
CASES["lv_h64_c128_b150_two_tiles"] = ("lv", 150, 12, dict(context_dim=128, hidden_dim=64, num_layers=2))
CASES["ou_h64_c128_l1_b130"] = ("ou", 130, 9, dict(context_dim=128, hidden_dim=64, num_layers=1))
CASES["l96s4_h64_c128_l2"] = ("l96", 5, 21, dict(context_dim=128, hidden_dim=64, num_layers=2, state_dim=4))
CASES["lv_h64_c128_t1"] = ("lv", 4, 1, dict(context_dim=128, hidden_dim=64, num_layers=2))   # single step: no carried state
CASES["ou_h64_c128_l1_t2"] = ("ou", 3, 2, dict(context_dim=128, hidden_dim=64, num_layers=1))
TC_OK = {"lv_h64_c128_l2", "ou_h64_c256_l1", "lv_h64_c256_l2", "l96s6_h64_c128_l2", "lv_h64_c128_b150_two_tiles",
         "ou_h64_c128_l1_b130", "l96s4_h64_c128_l2", "lv_h64_c128_t1", "ou_h64_c128_l1_t2"}
# wide-state tensor-core recurrence (4 < S <= 10, NL = 2; path_tcw.cu): ragged two-tile batch, odd S, single / two steps
CASES["l96s10_h64_c128_b150_tcw"] = ("l96", 150, 12, dict(context_dim=128, hidden_dim=64, num_layers=2, state_dim=10))
CASES["l96s5_h64_c128_tcw"] = ("l96", 5, 21, dict(context_dim=128, hidden_dim=64, num_layers=2, state_dim=5))
CASES["l96s7_h64_c256_t1_tcw"] = ("l96", 3, 1, dict(context_dim=256, hidden_dim=64, num_layers=2, state_dim=7))
CASES["l96s10_h64_c128_t2_tcw"] = ("l96", 130, 2, dict(context_dim=128, hidden_dim=64, num_layers=2, state_dim=10))
CASES["l96s9_h64_c128_tcw"] = ("l96", 4, 33, dict(context_dim=128, hidden_dim=64, num_layers=2, state_dim=9))
# staged ELBO kernels (wide state, user SDE, B >= 64): several 128-point chunks with a halo (T = 300), ragged last chunk
CASES["l96s6_h32_b70_t300_staged_elbo"] = ("l96", 70, 300, dict(context_dim=16, hidden_dim=32, num_layers=2, state_dim=6))
TCW = {"l96s10_h64_c128_b150_tcw", "l96s5_h64_c128_tcw", "l96s7_h64_c256_t1_tcw", "l96s10_h64_c128_t2_tcw", "l96s9_h64_c128_tcw"}
TC_OK |= TCW
TC_REC_OK = set(TC_OK)
NO_TC = 0x100  # VISDE_FLAG_NO_TENSOR_CORES
# batch sizes that do not divide the tile of the batch-tiled family
CASES["lv_h64_b37_ragged_tile"] = ("lv", 37, 9, dict(context_dim=16, hidden_dim=64, num_layers=2))
CASES["ou_h32_b13_ragged_tile"] = ("ou", 13, 7, dict(context_dim=8, hidden_dim=32, num_layers=1))
FAST_OK = {"lv_h64_b37_ragged_tile", "ou_h32_b13_ragged_tile", "ou_h32_l2", "lv_h16_l1", "ou_h64_l2", "lv_h64_l2", "lv_h48_l2", "ou_h20_l1", "l96s3_h64_l2",
           "lv_h64_c128_l2", "ou_h64_c256_l1", "lv_h64_c256_l2", "lv_h64_c128_b150_two_tiles", "ou_h64_c128_l1_b130",
           "l96s4_h64_c128_l2", "lv_h64_c128_t1", "ou_h64_c128_l1_t2"}


# batches above 4 x #SM take the 8-trajectory tiles of the batch-tiled family (forward and backward); ragged last tile
CASES["lv_h64_b601_tile8"] = ("lv", 601, 6, dict(context_dim=16, hidden_dim=64, num_layers=2))
CASES["ou_h32_l1_b610_tile8"] = ("ou", 610, 5, dict(context_dim=8, hidden_dim=32, num_layers=1))
CASES["l96s4_h64_b597_tile8"] = ("l96", 597, 4, dict(context_dim=8, hidden_dim=64, num_layers=2, state_dim=4))
FAST_OK |= {"lv_h64_b601_tile8", "ou_h32_l1_b610_tile8", "l96s4_h64_b597_tile8"}

# wide-state register-resident family (4 < S <= 16, H <= 64, NL <= 2)
CASES["l96s16_h32_l1"] = ("l96", 3, 9, dict(context_dim=8, hidden_dim=32, num_layers=1, state_dim=16))
CASES["l96s5_h64_l2_b150"] = ("l96", 150, 5, dict(context_dim=16, hidden_dim=64, num_layers=2, state_dim=5))
FASTS_OK = {"l96s10_h64_l2", "l96s6_h64_c128_l2", "l96s16_h32_l1", "l96s5_h64_l2_b150", "l96s6_h32_b70_t300_staged_elbo"} | TCW


def _variants(name):
    from viforsdes_b200 import _lib

    v = [_lib.VARIANT_GENERIC, _lib.VARIANT_FAST, _lib.VARIANT_TILED] if name in FAST_OK else [_lib.VARIANT_GENERIC]
    if name in FASTS_OK:
        v.append(_lib.VARIANT_FAST)
    if name in TC_OK:  # the same kernels with the GEMM stages forced onto the fp32 SIMT path
        v += [x | NO_TC for x in v]
    if name in TC_REC_OK:  # gate GEMMs of the recurrence on tcgen05 (fp16 hi/lo 3-pass split)
        v.append(_lib.VARIANT_TC)
    return v


@pytest.fixture(autouse=True)
def _reset_variant():
    from viforsdes_b200 import _lib, ops

    yield
    ops.set_variant(_lib.VARIANT_AUTO)


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_oracle(name):
    from viforsdes_b200 import ops

    kind, B, T, kw = CASES[name]
    p = O.make_problem(kind, B, T, **kw)
    ref = O.sample_paths(p.weights, p.x0, p.context, p.theta, p.eps, p.dt)
    for v in _variants(name):
        ops.set_variant(v)
        head = build_head(p).eval()
        x0, _, view, theta, eps = cuda_inputs(p)
        out = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
        for a, r, nm in zip(out, ref, ("paths", "means", "chol")):
            assert_close(a, r, name=f"{name}/v{v}/{nm}")
        assert torch.all(torch.triu(out[2], diagonal=1) == 0), "upper triangle of chol must be exactly zero"


@pytest.mark.parametrize("name", list(CASES))
def test_backward_matches_oracle_autograd(name):
    """All 4*NL+5 gradients for random upstream cotangents (gP, gM, gL)."""
    from viforsdes_b200 import ops

    kind, B, T, kw = CASES[name]
    p = O.make_problem(kind, B, T, **kw)
    g = torch.Generator().manual_seed(7)
    S = p.weights.state_dim
    gP, gM, gL = torch.randn(B, T + 1, S, generator=g), torch.randn(B, T, S, generator=g), torch.randn(B, T, S, S, generator=g)
    nl = p.weights.num_layers
    names = (["x0", "context", "theta"] + [f"w_ih_l{k}" for k in range(nl)] + [f"w_hh_l{k}" for k in range(nl)]
             + [f"b_ih_l{k}" for k in range(nl)] + [f"b_hh_l{k}" for k in range(nl)] + ["out_w", "out_b"])

    def oracle(dtype):
        w = p.weights.map(lambda t: t.to(dtype).clone().requires_grad_(True))
        x0 = p.x0.to(dtype).clone().requires_grad_(True)
        ctx = p.context.to(dtype).clone().requires_grad_(True)
        theta = p.theta.to(dtype).clone().requires_grad_(True)
        outs = O.sample_paths(w, x0, ctx, theta, p.eps.to(dtype), p.dt)
        gr = torch.autograd.grad(list(outs), [x0, ctx, theta, *w.tensors()], [gP.to(dtype), gM.to(dtype), gL.to(dtype)])
        return dict(zip(names, gr))

    ref, ref64 = oracle(torch.float32), oracle(torch.float64)
    for v in _variants(name):
        ops.set_variant(v)
        head = build_head(p)
        cx0, full, view, cth, ceps = cuda_inputs(p)
        out = head.sample_diffusion_paths(cx0, view, cth, ceps, p.dt)
        torch.autograd.backward(list(out), [gP.cuda(), gM.cuda(), gL.cuda()])
        got = {"x0": cx0.grad, "context": full.grad[:, :T], "theta": cth.grad, **head_grads(head)}
        assert torch.all(full.grad[:, T] == 0)
        for nm in names:
            assert_parity(got[nm], ref[nm], ref64[nm], name=f"{name}/v{v}/grad_{nm}")


@pytest.mark.parametrize("name", ["ou_h64_l2", "lv_h64_l2", "lv_h16_l1", "l96s4_h24_l3", "l96s10_h64_l2",
                                  "lv_h64_c128_l2", "ou_h64_c256_l1", "l96s6_h64_c128_l2", "l96s6_h32_b70_t300_staged_elbo",
                                  "l96s5_h64_l2_b150"])
def test_elbo_iteration_matches_oracle(name):
    """paths, ELBO terms and every gradient of -mean(obs + sde - gen + jac): the full hot path."""
    kind, B, T, kw = CASES[name]
    p = O.make_problem(kind, B, T, **kw)
    r32, r64 = oracle_refs(p)
    check_iteration(run_cuda_fwd_bwd(p), r32, r64, tag=f"{name}/")


@pytest.mark.parametrize("name", ["lv_h64_c128_b150_two_tiles", "ou_h64_c128_l1_b130", "l96s4_h64_c128_l2",
                                  "l96s10_h64_c128_b150_tcw", "l96s5_h64_c128_tcw", "l96s6_h64_c128_l2"])
def test_elbo_iteration_tensor_core_recurrence(name):
    """The same complete iteration with the recurrence FORCED onto the tcgen05 families (AUTO picks them only from
    B >= 3072): narrow (S <= 4) and wide-state (4 < S <= 10) kernels, ragged two-tile batches."""
    from viforsdes_b200 import _lib, ops

    kind, B, T, kw = CASES[name]
    p = O.make_problem(kind, B, T, **kw)
    r32, r64 = oracle_refs(p)
    ops.set_variant(_lib.VARIANT_TC)
    check_iteration(run_cuda_fwd_bwd(p), r32, r64, tag=f"tc/{name}/")


@pytest.mark.parametrize("name", ["stepwise_ou", "stepwise_lv", "stepwise_l96", "stepwise_ou_h64"])
def test_matches_reference_golden(name):
    """Directly against outputs of the real reference (tests/golden/make_golden.py)."""
    g = torch.load(GOLD / f"{name}.pt")
    p = O.make_problem(g["kind"], g["B"], g["T"], **g["kw"])
    paths, means, chol, terms, grads = run_cuda_fwd_bwd(p)
    r64 = O.run_fwd_bwd(p, dtype=torch.float64)
    assert_parity(paths, g["paths"], r64[0], name="paths")
    assert_parity(means, g["means"], r64[1], name="means")
    assert_parity(chol, g["chol"], r64[2], name="chol")
    for j, nm in enumerate(("obs", "sde", "gen", "jac")):
        tol = 2e-3 if nm == "jac" else 2e-4  # jac_mean is recovered by subtraction in the golden script
        assert abs(terms[:, j].mean().item() - g[f"{nm}_mean"].item()) <= tol * max(1.0, abs(g[f"{nm}_mean"].item()))
    ref = {"x0": g["g_x0"], "context": g["g_context"], "theta": g["g_theta"], "out_w": g["g_out_w"], "out_b": g["g_out_b"]}
    for k in range(p.weights.num_layers):
        ref[f"w_ih_l{k}"], ref[f"w_hh_l{k}"] = g[f"g_weight_ih_l{k}"], g[f"g_weight_hh_l{k}"]
        ref[f"b_ih_l{k}"], ref[f"b_hh_l{k}"] = g[f"g_bias_ih_l{k}"], g[f"g_bias_hh_l{k}"]
    for nm, r in ref.items():
        assert_parity(grads[nm], r, r64[4][nm], name=f"grad_{nm}")


@pytest.mark.parametrize("name", ["triton_lv", "triton_l96"])
def test_matches_reference_triton_golden(name):
    """Same cotangents as the reference's own fused kernels were run with (interpreter mode)."""
    g = torch.load(GOLD / f"{name}.pt")
    p = O.make_problem(g["kind"], g["B"], g["T"], **g["kw"])
    nl = p.weights.num_layers
    d = torch.float64
    w64 = p.weights.map(lambda t: t.to(d).clone().requires_grad_(True))
    lx0, lctx, lth = (t.to(d).clone().requires_grad_(True) for t in (p.x0, p.context, p.theta))
    o64 = O.sample_paths(w64, lx0, lctx, lth, p.eps.to(d), p.dt)
    g64 = torch.autograd.grad(list(o64), [lx0, lctx, lth, *w64.tensors()], [g["gP"].to(d), g["gM"].to(d), g["gL"].to(d)])
    w_ih64, w_hh64, b_ih64, b_hh64 = (g64[3 + i * nl:3 + (i + 1) * nl] for i in range(4))

    head = build_head(p)
    x0, full, view, theta, eps = cuda_inputs(p)
    out = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
    for a, r64, nm in zip(out, o64, ("paths", "means", "chol")):
        assert_parity(a, g[nm], r64, name=nm)
    torch.autograd.backward(list(out), [g["gP"].cuda(), g["gM"].cuda(), g["gL"].cuda()])
    hg = head_grads(head)
    assert_parity(x0.grad, g["g_x0"], g64[0], name="g_x0")
    assert_parity(full.grad[:, :g["T"]], g["g_context"], g64[1], name="g_context")
    assert_parity(theta.grad, g["g_theta"], g64[2], name="g_theta")
    assert_parity(hg["w_ih_l0"], g["g_w_ih_l0"], w_ih64[0], name="g_w_ih_l0")
    assert_parity(hg["w_hh_l0"], g["g_w_hh_l0"], w_hh64[0], name="g_w_hh_l0")
    assert_parity(hg["b_ih_l0"], g["g_b_ih_l0"], b_ih64[0], name="g_b_ih_l0")
    assert_parity(hg["b_hh_l0"], g["g_b_hh_l0"], b_hh64[0], name="g_b_hh_l0")
    if nl > 1:
        for nm, key, r in (("w_ih", "g_w_ih_stack", w_ih64), ("w_hh", "g_w_hh_stack", w_hh64),
                           ("b_ih", "g_b_ih_stack", b_ih64), ("b_hh", "g_b_hh_stack", b_hh64)):
            assert_parity(torch.stack([hg[f"{nm}_l{k}"] for k in range(1, nl)]), g[key], torch.stack(list(r[1:])), name=key)
    assert_parity(hg["out_w"], g["g_out_w"], g64[3 + 4 * nl], name="g_out_w")
    assert_parity(hg["out_b"], g["g_out_b"], g64[4 + 4 * nl], name="g_out_b")


def test_diag_floor_branch_and_gradient_rule():
    """Cholesky diagonal floored at DIAG_MIN; gradient passes iff raw >= bound or grad < 0
    (primitives/bounds.py:20).  A bias of -0.5 on the diagonal rows forces the clamped branch."""
    from viforsdes_b200 import _lib, ops

    p = O.make_problem("lv", 4, 15, context_dim=8, hidden_dim=32, num_layers=2)
    S = 2
    for d in range(S):
        p.weights.out_b[S + d * (d + 3) // 2] = -0.5 if d == 0 else 0.02
    ref = O.sample_paths(p.weights, p.x0, p.context, p.theta, p.eps, p.dt)
    assert (ref[2][:, :, 0, 0] == O.DIAG_MIN).any(), "test must exercise the clamped branch"
    r32, r64 = oracle_refs(p)
    for v in (_lib.VARIANT_GENERIC, _lib.VARIANT_FAST):
        ops.set_variant(v)
        # the floor's gradient rule is DISCONTINUOUS in raw (bias 0.02 puts L_11 within rounding distance of the floor), so
        # the fp32 and fp64 oracles themselves take different branches on a few entries: wider noise cap for this test only
        check_iteration(run_cuda_fwd_bwd(p), r32, r64, tag=f"diagfloor/v{v}/", noise_cap=2e-2)


def test_bf16_and_contiguous_context():
    """AMP path: bf16 context is consumed in place, grad_context comes back bf16."""
    p = O.make_problem("ou", 4, 20, context_dim=32, hidden_dim=64, num_layers=2)
    pb = O.make_problem("ou", 4, 20, context_dim=32, hidden_dim=64, num_layers=2)
    pb.context = p.context.to(torch.bfloat16).to(torch.float32)  # oracle sees the rounded values
    r_paths, _, _, r_terms, r_grads = O.run_fwd_bwd(pb)
    r64 = O.run_fwd_bwd(pb, dtype=torch.float64)
    paths, _, _, terms, grads = run_cuda_fwd_bwd(p, ctx_dtype=torch.bfloat16)
    assert grads["context"].dtype == torch.bfloat16
    assert_close(paths, r_paths, name="paths")
    assert_close(grads["context"].float(), r_grads["context"], rtol=1e-2, atol_scale=1e-2, name="grad_context(bf16)")
    for nm in ("x0", "theta", "w_ih_l0", "w_hh_l1", "out_w"):
        assert_parity(grads[nm], r_grads[nm], r64[4][nm], name=nm)
    # contiguous [B,T,C] fp32 input gives bit-identical results to the strided view
    a = run_cuda_fwd_bwd(p, strided=True)
    b = run_cuda_fwd_bwd(p, strided=False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[4]["w_ih_l0"], b[4]["w_ih_l0"])


def test_bf16_context_takes_the_tensor_core_gemms():
    """bf16 context on shapes eligible for the tcgen05 GEMM stages (what the reference's autocast encoder feeds the
    head): widened once to fp32 in the workspace, then the same tensor-core K0 / K3 / K4 as for fp32 context -- both
    with the register-resident and with the tensor-core recurrence family.  Same numbers as the fp32 SIMT stages on
    the bf16-rounded values."""
    from viforsdes_b200 import _lib, ops

    p = O.make_problem("lv", 130, 21, context_dim=128, hidden_dim=64, num_layers=2)
    pb = O.make_problem("lv", 130, 21, context_dim=128, hidden_dim=64, num_layers=2)
    pb.context = p.context.to(torch.bfloat16).to(torch.float32)
    r32, r64 = oracle_refs(pb)
    for v in (_lib.VARIANT_AUTO, _lib.VARIANT_TC, _lib.VARIANT_FAST | NO_TC):
        ops.set_variant(v)
        paths, means, chol, terms, grads = run_cuda_fwd_bwd(p, ctx_dtype=torch.bfloat16)
        assert grads["context"].dtype == torch.bfloat16
        for a, b32, b64, nm in zip((paths, means, chol), r32[:3], r64[:3], ("paths", "means", "chol")):
            assert_parity(a, b32, b64, name=f"v{v}/{nm}")
        assert_close(grads["context"].float(), r64[4]["context"], rtol=1e-2, atol_scale=1e-2, name=f"v{v}/grad_context(bf16)")
        for nm in r64[4]:
            if nm != "context":
                assert_parity(grads[nm], r32[4][nm], r64[4][nm], name=f"v{v}/grad_{nm}")


def test_edge_shapes():
    """B = 1, T = 1, and empty inputs (B = 0 / T = 0) do not crash and match the oracle."""
    for B, T in ((1, 1), (1, 5), (3, 1)):
        p = O.make_problem("lv", B, T, context_dim=8, hidden_dim=16, num_layers=2)
        r32, r64 = oracle_refs(p)
        check_iteration(run_cuda_fwd_bwd(p), r32, r64, tag=f"B{B} T{T} ")
    p = O.make_problem("ou", 2, 4, context_dim=8, hidden_dim=16, num_layers=1)
    head = build_head(p).eval()
    z = torch.zeros
    out = head.sample_diffusion_paths(z(0, 1).cuda(), z(0, 4, 8).cuda(), z(0, 3).cuda(), z(0, 4, 1).cuda(), 0.05)
    assert out[0].shape == (0, 5, 1) and out[2].shape == (0, 4, 1, 1)
    out = head.sample_diffusion_paths(p.x0.cuda(), z(2, 0, 8).cuda(), p.theta.cuda(), z(2, 0, 1).cuda(), 0.05)
    assert out[0].shape == (2, 1, 1) and torch.equal(out[0][:, 0].cpu(), p.x0)


def test_bad_config_raises_value_error():
    """Error behaviour of the reference operator (models/head.py:33-36): ValueError."""
    from viforsdes_b200.head import DiffusionTransitionHead, HeadConfig

    with pytest.raises(ValueError):
        DiffusionTransitionHead(1, 8, 3, HeadConfig(hidden_dim=16, num_layers=5))
    head = DiffusionTransitionHead(1, 8, 3, HeadConfig(hidden_dim=16, num_layers=1)).cuda()
    with pytest.raises(ValueError):
        head.sample_diffusion_paths(torch.zeros(2, 1).cuda(), torch.zeros(2, 4, 9).cuda(), torch.zeros(2, 3).cuda(),
                                    torch.zeros(2, 4, 1).cuda(), 0.05)
    with pytest.raises(ValueError):
        head.sample_diffusion_paths(torch.zeros(2, 1).cuda(), torch.zeros(2, 4, 8).cuda(), torch.zeros(2, 3).cuda(),
                                    torch.zeros(2, 4, 1).cuda(), -1.0)


def test_session_host_step_matches_oracle():
    """The host-buffer C-ABI entry (H2D -> fwd -> ELBO -> bwd -> D2H) used by bench.py's e2e leg."""
    from viforsdes_b200.session import HostSession

    p = O.make_problem("lv", 6, 30, context_dim=32, hidden_dim=64, num_layers=2)
    r32, r64 = oracle_refs(p)
    sess = HostSession.from_problem(p, want_grad_context=True)
    res = sess.step()
    for j, nm in enumerate(("obs", "sde", "gen", "jac")):
        assert_parity(res["terms"][:, j], getattr(r32[3], nm), getattr(r64[3], nm), name=nm)
    for nm in r64[4]:
        assert_parity(res["grads"][nm], r32[4][nm], r64[4][nm], name=f"grad_{nm}")
    assert sess.h2d_bytes > 0 and sess.d2h_bytes > 0 and sess.launches > 0
    sess.close()


def test_session_pipelined_submit_wait_matches_sync():
    """visde_session_submit / _wait (two iterations in flight, H2D of i+1 overlapping the kernels of i)
    returns bit-identical outputs to the synchronous visde_session_step, in submission order, and
    refuses a third in-flight iteration."""
    from viforsdes_b200.session import HostSession

    p = O.make_problem("lv", 6, 30, context_dim=32, hidden_dim=64, num_layers=2)
    sess = HostSession.from_problem(p, want_grad_context=True)
    ref = sess.step()
    ref = {"terms": ref["terms"].clone(), "grads": {k: v.clone() for k, v in ref["grads"].items()}}
    sess.submit()
    sess.submit()
    with pytest.raises(ValueError):
        sess.submit()
    for _ in range(2):
        res = sess.wait()
        assert torch.equal(res["terms"], ref["terms"])
        for k, v in ref["grads"].items():
            assert torch.equal(res["grads"][k], v), k
    with pytest.raises(ValueError):
        sess.wait()
    # changed inputs between iterations are picked up by the second input set
    sess.submit()
    sess.wait()
    sess.eps.mul_(0.5)
    sess.submit()
    res2 = sess.wait()
    assert not torch.equal(res2["terms"], ref["terms"])
    sess.eps.mul_(2.0)
    assert torch.equal(sess.step()["terms"], ref["terms"])
    sess.close()


def test_full_size_properties():
    """BASELINE config 2 size (LV, B=128, T=800, C=256, H=64x2): size-independent properties --
    run-to-run determinism (no atomics), agreement of the two kernel families, linearity of the
    backward in its cotangents, and fp32 agreement with the fp64 oracle on a batch slice."""
    from viforsdes_b200 import _lib, ops

    p = O.make_problem("lv", 128, 800, context_dim=256, hidden_dim=64, num_layers=2)
    head = build_head(p)

    def run(scale=1.0):
        for q in head.parameters():
            q.grad = None
        x0, full, view, theta, eps = cuda_inputs(p)
        out = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
        g = torch.Generator(device="cuda").manual_seed(3)
        cts = [torch.randn(o.shape, device="cuda", generator=g) * scale for o in out]
        torch.autograd.backward(list(out), cts)
        return [o.detach() for o in out], {"x0": x0.grad, "context": full.grad, "theta": theta.grad, **head_grads(head)}

    o1, g1 = run()
    o2, g2 = run()
    for a, b in zip(o1, o2):
        assert torch.equal(a, b), "forward must be bit-deterministic"
    for k in g1:
        assert torch.equal(g1[k], g2[k]), f"backward must be bit-deterministic ({k})"
    _, g3 = run(scale=2.0)
    for k in g1:
        assert normwise(g3[k], 2.0 * g1[k]) < 1e-5, f"backward must be linear in the cotangents ({k})"
    ops.set_variant(_lib.VARIANT_GENERIC)
    o4, g4 = run()
    for a, b, nm in zip(o1, o4, ("paths", "means", "chol")):
        assert normwise(a, b) < 1e-4, f"fast vs generic {nm}"
    for k in g1:
        assert normwise(g1[k], g4[k]) < 2e-4, f"fast vs generic grad {k}: {normwise(g1[k], g4[k])}"
    ops.set_variant(_lib.VARIANT_AUTO)
    # fp64 oracle on the first 2 trajectories (forward only; seconds on CPU)
    p64 = O.make_problem("lv", 128, 800, context_dim=256, hidden_dim=64, num_layers=2)
    w64 = p64.weights.map(lambda t: t.double())
    ref = O.sample_paths(w64, p64.x0[:2].double(), p64.context[:2].double(), p64.theta[:2].double(),
                         p64.eps[:2].double(), p64.dt)
    for a, r, nm in zip(o1, ref, ("paths", "means", "chol")):
        assert normwise(a[:2], r) < 1e-4, f"{nm} vs fp64 oracle: {normwise(a[:2], r)}"


def test_tc_family_large_batch_properties():
    """Large batch (B = 3 200 = 25 tiles of 128 trajectories): AUTO must pick the tensor-core recurrence
    family; size-independent properties -- bit-determinism (fixed-order reductions, no atomics), agreement
    with the FP32 register-resident family on every output and gradient, linearity of the backward, and
    fp64-oracle agreement on a slice of the batch."""
    from viforsdes_b200 import _lib, ops

    p = O.make_problem("lv", 3200, 40, context_dim=128, hidden_dim=64, num_layers=2)
    head = build_head(p)

    def run(scale=1.0):
        for q in head.parameters():
            q.grad = None
        x0, full, view, theta, eps = cuda_inputs(p)
        out = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
        g = torch.Generator(device="cuda").manual_seed(3)
        cts = [torch.randn(o.shape, device="cuda", generator=g) * scale for o in out]
        torch.autograd.backward(list(out), cts)
        return [o.detach() for o in out], {"x0": x0.grad, "context": full.grad, "theta": theta.grad, **head_grads(head)}

    ops.set_variant(_lib.VARIANT_AUTO)
    o1, g1 = run()
    o2, g2 = run()
    for a, b in zip(o1, o2):
        assert torch.equal(a, b), "forward must be bit-deterministic"
    for k in g1:
        assert torch.equal(g1[k], g2[k]), f"backward must be bit-deterministic ({k})"
    _, g3 = run(scale=2.0)
    for k in g1:
        assert normwise(g3[k], 2.0 * g1[k]) < 1e-5, f"backward must be linear in the cotangents ({k})"
    ops.set_variant(_lib.VARIANT_TC)
    o5, _ = run()
    for a, b in zip(o1, o5):
        assert torch.equal(a, b), "AUTO at B = 3200 must be the tensor-core family"
    ops.set_variant(_lib.VARIANT_FAST)
    o4, g4 = run()
    for a, b, nm in zip(o1, o4, ("paths", "means", "chol")):
        assert normwise(a, b) < 1e-4, f"tc vs fast {nm}: {normwise(a, b)}"
    for k in g1:
        assert normwise(g1[k], g4[k]) < 2e-4, f"tc vs fast grad {k}: {normwise(g1[k], g4[k])}"
    ops.set_variant(_lib.VARIANT_AUTO)
    w64 = p.weights.map(lambda t: t.double())
    sl = slice(3197, 3200)  # last rows of the last (full) tile
    ref = O.sample_paths(w64, p.x0[sl].double(), p.context[sl].double(), p.theta[sl].double(), p.eps[sl].double(), p.dt)
    for a, r, nm in zip(o1, ref, ("paths", "means", "chol")):
        assert normwise(a[sl], r) < 1e-4, f"{nm} vs fp64 oracle: {normwise(a[sl], r)}"


def test_wide_tensor_core_recurrence_128_row_form_and_tile_loop():
    """149 tiles: more than SMs / 2, so the wide family runs its 128-row form (MMA M = 128, two threads per row) instead of the 64-row
    one every smaller case takes, and more than 148, so one CTA walks two tiles (barrier phases carried across tiles); ragged last tile."""
    from viforsdes_b200 import _lib, ops

    p = O.make_problem("l96", 148 * 128 + 37, 3, context_dim=128, hidden_dim=64, num_layers=2, state_dim=10)
    r32, r64 = oracle_refs(p)
    ops.set_variant(_lib.VARIANT_TC)
    check_iteration(run_cuda_fwd_bwd(p), r32, r64, tag="tc/l96s10_b18981_m128/")
