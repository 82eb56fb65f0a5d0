"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/visde.h declares, the ctypes prototypes cover them, and argument validation
returns the documented error codes (no compute is launched here)."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from viforsdes_b200 import _lib

    if not _lib.LIB_PATH.exists():
        _lib.build()
    return _lib.load()


def _declared_symbols():
    text = (ROOT / "include" / "visde.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(visde_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    from viforsdes_b200 import _lib

    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/visde.h but not exported by libvisde.so"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(syms)


def test_struct_layout_matches_header():
    from viforsdes_b200 import _lib

    assert C.sizeof(_lib.Dims) == 40           # 2 x int64 + 6 x int32
    assert C.sizeof(_lib.Weights) == 8 * 18    # 4 x 4 + 2 pointers
    assert C.sizeof(_lib.CtxView) == 32
    assert C.sizeof(_lib.Obs) == 40


def test_sizes_and_validation(lib):
    from viforsdes_b200 import _lib

    d = _lib.Dims(128, 800, 2, 256, 3, 64, 2, 0)
    stash = lib.visde_stash_bytes(C.byref(d))
    assert stash >= 128 * 800 * (2 * 5 * 64 + 3) * 4
    assert lib.visde_workspace_bytes(C.byref(d), 0) >= 128 * 800 * 192 * 4
    assert lib.visde_workspace_bytes(C.byref(d), 1) >= 128 * 800 * 2 * 4 * 64 * 4
    bad = _lib.Dims(1, 1, 1, 8, 3, 16, 5, 0)  # num_layers > MAX_LAYERS (kernels/constants.py:13)
    assert lib.visde_stash_bytes(C.byref(bad)) == 0
    assert b"num_layers" in lib.visde_last_error()
    w = _lib.Weights()
    rc = lib.visde_path_fwd(C.byref(bad), 0.1, None, None, None, None, C.byref(w), None, None, None, None, None, 0, None)
    assert rc == _lib.EINVAL
    with pytest.raises(ValueError):
        _lib.check(rc)
    ok = _lib.Dims(0, 4, 1, 8, 3, 16, 1, 0)  # empty batch: nothing to launch, success
    for k in range(1):
        w.w_ih[k] = w.w_hh[k] = w.b_ih[k] = w.b_hh[k] = 1
    w.out_w = w.out_b = 1
    assert lib.visde_path_fwd(C.byref(ok), 0.1, None, None, None, None, C.byref(w), None, None, None, None, None, 0, None) == 0


def test_caller_entry_points_validate_without_a_gpu(lib):
    """Argument validation of the SURVEY 8f entry points returns the documented codes before anything is launched."""
    from viforsdes_b200 import _lib

    # core/euler_maruyama.py:20-23 raises ValueError on dt <= 0; user SDEs have no device functor
    assert lib.visde_em_fwd(4, 4, _lib.SDE_GENERIC, 0, 0.05, None, None, None, 0, None, None) == _lib.EINVAL
    assert b"PyTorch" in lib.visde_last_error()
    assert lib.visde_em_fwd(4, 4, _lib.SDE_OU, 0, 0.0, None, None, None, 0, None, None) == _lib.EINVAL
    assert lib.visde_em_fwd(4, 4, _lib.SDE_LV, 3, 0.05, None, None, None, 0, None, None) == _lib.EINVAL  # NULL tensors
    assert lib.visde_em_fwd(0, 4, _lib.SDE_LV, 3, 0.05, None, None, None, 0, None, None) == _lib.OK       # empty batch
    assert lib.visde_em_bwd(0, 4, _lib.SDE_LV, 3, 0.05, None, None, 0, None, None, None, None, None) == _lib.OK
    assert lib.visde_em_bwd(2, 4, _lib.SDE_LV, 3, 0.05, None, None, 0, None, None, None, None, None) == _lib.EINVAL
    assert lib.visde_philox_normal(1, 2, 2, 17, None, None) == _lib.EINVAL  # S <= VISDE_MAX_STATE
    assert lib.visde_philox_normal(1, 0, 2, 2, None, None) == _lib.OK
    # posterior summary
    assert lib.visde_path_summary_workspace_bytes(1000, 801, 2) >= 3 * 801 * 2 * 4
    assert lib.visde_path_summary(10, 5, 0, 0, None, None, None, None, None, 0, None) == _lib.EINVAL
    assert lib.visde_path_summary(10, 0, 2, 0, None, None, None, None, None, 0, None) == _lib.OK  # no grid points
    assert lib.visde_path_summary(10, 5, 2, 0, 1, None, 1, 1, None, 0, None) == _lib.EWORKSPACE
    # optimiser tail
    assert lib.visde_grad_sqnorm_workspace_bytes() >= 4 * 148
    assert lib.visde_grad_sqnorm(16, 16, None, 0, None, None, None, 0, None) == _lib.EINVAL           # sqnorm NULL
    assert lib.visde_grad_sqnorm(16, 16, None, 0, 16, None, None, 0, None) == _lib.EWORKSPACE
    assert lib.visde_grad_sqnorm(16, 4, None, 0, 16, None, 16, 1 << 20, None) == _lib.EINVAL          # unaligned gradient pointer
    args = (0.001, 0.9, 0.999, 1e-8, 0.01)
    assert lib.visde_adamw_ema_step(0, None, None, None, None, None, *args, 1, 1.0, None, None, 0.999, None, None) == _lib.OK
    assert lib.visde_adamw_ema_step(8, 16, 16, 16, 16, None, *args, 0, 1.0, None, None, 0.999, None, None) == _lib.EINVAL  # step from 1
    assert lib.visde_adamw_ema_step(8, 16, 16, 16, 16, None, 0.001, 1.0, 0.999, 1e-8, 0.01, 1, 1.0, None, None, 0.999,
                                    None, None) == _lib.EINVAL  # beta1 must be < 1
    assert lib.visde_adamw_ema_step(8, None, 16, 16, 16, None, *args, 1, 1.0, None, None, 0.999, None, None) == _lib.EINVAL


def test_host_mirrors_of_the_callers_validate_on_cpu():
    """euler_maruyama keeps the reference's ValueErrors and its PyTorch loop for user SDEs (CPU tensors allowed there);
    the fused ops refuse CPU tensors; FlatParameters aliases parameters and gradients into flat buffers."""
    import torch
    from torch import nn

    from viforsdes_b200 import sde as vs
    from viforsdes_b200.euler_maruyama import euler_maruyama, pretrain_mse
    from viforsdes_b200.optim import FlatParameters, FusedAdamWEma
    from viforsdes_b200.posterior import summarise_paths
    from viforsdes_b200.state_space import StateSpace

    ou = vs.OrnsteinUhlenbeck()
    x0, th = torch.zeros(3, 1), torch.tensor([[1.0, 0.5, 0.2]]).repeat(3, 1)
    with pytest.raises(ValueError, match="dt must be positive"):
        euler_maruyama(ou, x0, th, 1.0, 0.0)
    with pytest.raises(ValueError, match="time_horizon must be positive"):
        euler_maruyama(ou, x0, th, 0.0, 0.1)
    # CPU tensors: the reference loop (a device functor exists only on CUDA)
    noise = torch.randn(3, 10, 1, generator=torch.Generator().manual_seed(0))
    paths = euler_maruyama(ou, x0, th, 1.0, 0.1, noise=noise)
    x = x0.clone()
    for t in range(10):
        x = x + th[:, 0:1] * (th[:, 1:2] - x) * 0.1 + th[:, 2:3] * noise[:, t] * 0.1**0.5
    assert paths.shape == (3, 11, 1) and torch.allclose(paths[:, -1], x, atol=1e-6)
    mse = pretrain_mse(ou, th, torch.tensor([0.0, 1.0]), torch.tensor([[0.0], [0.3]]), 1.0, 0.1, noise=noise)
    assert torch.isfinite(mse)
    with pytest.raises(RuntimeError, match="CUDA"):
        torch.ops.visde.em_fwd(x0, th, None, 0, 4, 0.1, ou.device_kind, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        summarise_paths(torch.zeros(4, 3, 2), StateSpace(2, [0]))
    lin = nn.Linear(3, 2)
    w0 = lin.weight.detach().clone()
    flat = FlatParameters([[lin.weight], [lin.bias]])
    assert flat.segments == [(0, 8), (8, 12)] and torch.equal(lin.weight, w0)
    lin(torch.ones(1, 3)).sum().backward()
    assert flat.grads[:6].abs().sum() > 0 and lin.weight.grad.data_ptr() == flat.grads.data_ptr()
    flat.zero_grad()
    assert lin.weight.grad.abs().sum() == 0
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedAdamWEma(flat, lrs=[1e-3, 1e-3])
    with pytest.raises(ValueError):
        FusedAdamWEma(flat, lrs=[1e-3])


def test_ops_reject_cpu_tensors():
    """No CPU fallback: the product path fails loudly off-GPU."""
    import torch

    from viforsdes_b200.head import DiffusionTransitionHead, HeadConfig

    head = DiffusionTransitionHead(1, 8, 3, HeadConfig(16, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        head.sample_diffusion_paths(torch.zeros(2, 1), torch.zeros(2, 4, 8), torch.zeros(2, 3), torch.zeros(2, 4, 1), 0.05)


def test_checkpoint_keys_match_reference():
    """state_dict keys of the head are the reference's so posterior checkpoints load."""
    from viforsdes_b200.head import DiffusionTransitionHead, HeadConfig

    keys = set(DiffusionTransitionHead(2, 16, 3, HeadConfig(32, 2)).state_dict())
    want = {"_tril_rows", "_tril_cols", "_diag_mask", "out_proj.weight", "out_proj.bias"}
    for k in range(2):
        want |= {f"gru.weight_ih_l{k}", f"gru.weight_hh_l{k}", f"gru.bias_ih_l{k}", f"gru.bias_hh_l{k}"}
    assert keys == want


def test_device_adam_matches_torch_adam_and_skips():
    """The sync-free Adam of the pre-training loop == torch.optim.Adam + clip_grad_norm_(1.0), and a masked step leaves
    parameter and state untouched (what the reference does with `if torch.isfinite(mse)`, inference/trainer.py:238-241)."""
    import torch

    from viforsdes_b200.euler_maruyama import PretrainConfig, _DeviceAdam

    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(6, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=0.02)
    mine = p0.clone().requires_grad_(True)
    dev = _DeviceAdam(mine, 0.02)
    for k in range(8):
        grad = torch.randn(6, generator=g) * (5.0 if k % 2 else 0.1)
        if k == 3:  # skipped step (non-finite objective)
            before = mine.detach().clone()
            dev.step(torch.full((6,), float("nan")), torch.tensor(False))
            assert torch.equal(mine.detach(), before)
            continue
        ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        dev.step(grad, torch.tensor(True))
    assert torch.allclose(mine.detach(), ref.detach(), rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError):
        PretrainConfig(n_iterations=0)


def test_prototype_arity_matches_header():
    """Every ctypes prototype has exactly as many arguments as the C declaration in include/visde.h (ABI drift guard)."""
    from viforsdes_b200 import _lib

    text = (ROOT / "include" / "visde.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = dict(re.findall(r"\b(visde_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text))
    assert set(decls) == set(_lib.PROTOTYPES)
    for name, params in decls.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.PROTOTYPES[name][1]), f"{name}: header declares {n} parameters, ctypes binds {len(_lib.PROTOTYPES[name][1])}"


def test_auto_family_selection_table(lib):
    """visde_recurrence_family pins what AUTO picks (148 SMs: the default when no device is present, and a B200's count):
    the measured waves x cost model of tiled_batch_tile (profiles/r1_ou_batch_sweep.md) and the family boundaries."""
    from viforsdes_b200 import _lib as L

    def fam(B, back, S=1, H=64, NL=2, Cd=256, variant=L.VARIANT_AUTO, T=100):
        d = L.Dims(B, T, S, Cd, 3, H, NL, variant)
        return lib.visde_recurrence_family(C.byref(d), back)

    F, T4, T8, TC, FS, G = L.FAMILY_FAST, L.FAMILY_TILED4, L.FAMILY_TILED8, L.FAMILY_TC, L.FAMILY_FAST_S, L.FAMILY_GENERIC
    #            B: (forward, backward)
    table = {128: (F, F), 148: (F, F), 200: (F, F), 296: (F, F), 444: (T4, F), 592: (T4, T4), 740: (T8, F),
             1024: (T8, T8), 1184: (T8, T8), 2048: (T8, T8), 3071: (T8, T8), 3072: (TC, TC), 65536: (TC, TC)}
    for B, (fw, bw) in table.items():
        assert (fam(B, 0), fam(B, 1)) == (fw, bw), f"B={B}: {(fam(B, 0), fam(B, 1))}"
    # wide state (BASELINE config 5, S = 10): the wide tensor-core family from 1 024 trajectories per GPU (every shard size of the
    # strong-scaled run, N = 1 .. 8), the register-resident wide family below (measured crossover B ~ 900, profiles/r2_config5.md)
    assert fam(8192, 0, S=10) == TC and fam(8192, 1, S=10) == TC and fam(4096, 1, S=10) == TC and fam(2048, 0, S=10) == TC
    assert fam(1536, 1, S=10) == TC and fam(1024, 0, S=10) == TC and fam(1023, 1, S=10) == FS and fam(768, 0, S=10) == FS
    assert fam(8192, 0, S=10, NL=1) == FS and fam(8192, 0, S=12) == FS    # outside the wide tensor-core shapes (NL = 2, S <= 10)
    assert fam(8192, 0, H=128) == G and fam(128, 1, NL=3) == G          # outside the register-resident shapes
    assert fam(8192, 0, Cd=64) == T8                                      # no tcgen05 K0 for this context width: no TC recurrence
    assert fam(8192, 0, variant=L.VARIANT_FAST) == F and fam(37, 1, variant=L.VARIANT_TILED) == T4
    assert fam(601, 0, variant=L.VARIANT_TILED) == T8 and fam(130, 0, variant=L.VARIANT_TC, Cd=128) == TC
    assert fam(8192, 0, variant=L.VARIANT_AUTO | 0x100) == T8            # VISDE_FLAG_NO_TENSOR_CORES
    assert fam(1, 0, NL=5) == L.EINVAL
