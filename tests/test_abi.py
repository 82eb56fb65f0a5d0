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
