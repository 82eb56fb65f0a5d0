"""ctypes binding of libvisde.so (include/visde.h).  The library is built in-tree by
``make -C viforsdes_b200/csrc`` (``__graft_entry__.build()``); there is no CPU fallback:
``load()`` raises if the shared library is missing."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libvisde.so"
MAX_LAYERS = 4
VERSION = 2  # include/visde.h VISDE_VERSION

F32, BF16 = 0, 1
SDE_GENERIC, SDE_OU, SDE_LV = 0, 1, 2
VARIANT_AUTO, VARIANT_GENERIC, VARIANT_FAST, VARIANT_TILED, VARIANT_TC = 0, 1, 2, 3, 4
OK, EINVAL, ECUDA, EWORKSPACE = 0, -1, -2, -3
FAMILY_GENERIC, FAMILY_FAST, FAMILY_TILED4, FAMILY_TILED8, FAMILY_TC, FAMILY_FAST_S = range(6)
FAMILY_NAMES = ("generic", "fast", "tiled4", "tiled8", "tc", "fast_s")
STAGES = ("K0_ctx_gemm", "K1_path_fwd", "K5_elbo_fwd", "K6_elbo_bwd", "K2_path_bwd", "K3_grad_ctx", "K4_wgrad")

_fp = C.c_void_p


class Dims(C.Structure):
    _fields_ = [("B", C.c_int64), ("T", C.c_int64), ("S", C.c_int32), ("C", C.c_int32), ("P", C.c_int32),
                ("H", C.c_int32), ("NL", C.c_int32), ("variant", C.c_int32)]


class Weights(C.Structure):
    _fields_ = [("w_ih", _fp * MAX_LAYERS), ("w_hh", _fp * MAX_LAYERS), ("b_ih", _fp * MAX_LAYERS),
                ("b_hh", _fp * MAX_LAYERS), ("out_w", _fp), ("out_b", _fp)]


class CtxView(C.Structure):
    _fields_ = [("ptr", _fp), ("batch_stride", C.c_int64), ("time_stride", C.c_int64), ("dtype", C.c_int32)]


SDE_EVAL_FN = C.CFUNCTYPE(C.c_int, _fp, _fp, _fp, _fp, _fp, _fp)
SDE_VJP_FN = C.CFUNCTYPE(C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp)


class UserSde(C.Structure):
    _fields_ = [("eval", SDE_EVAL_FN), ("vjp", SDE_VJP_FN), ("user", _fp)]


class Obs(C.Structure):
    _fields_ = [("n_obs", C.c_int32), ("obs_dim", C.c_int32), ("idx", _fp), ("values", _fp),
                ("obs_matrix", _fp), ("variance", C.c_float)]


# name -> (restype, argtypes); kept in sync with include/visde.h (tests/test_abi.py parses the header)
PROTOTYPES = {
    "visde_version": (C.c_int, []),
    "visde_last_error": (C.c_char_p, []),
    "visde_recurrence_family": (C.c_int, [C.POINTER(Dims), C.c_int]),
    "visde_stash_bytes": (C.c_size_t, [C.POINTER(Dims)]),
    "visde_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int]),
    "visde_path_fwd": (C.c_int, [C.POINTER(Dims), C.c_float, _fp, C.POINTER(CtxView), _fp, _fp, C.POINTER(Weights),
                                 _fp, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "visde_path_bwd": (C.c_int, [C.POINTER(Dims), C.c_float, _fp, _fp, _fp, C.POINTER(CtxView), _fp, _fp,
                                 C.POINTER(Weights), _fp, _fp, _fp, C.POINTER(CtxView), _fp, C.POINTER(Weights),
                                 _fp, C.c_size_t, _fp]),
    "visde_elbo_fwd": (C.c_int, [C.POINTER(Dims), C.c_float, C.c_int, C.c_uint32, _fp, _fp, _fp, _fp, _fp, _fp,
                                 C.POINTER(Obs), _fp, _fp]),
    "visde_elbo_bwd": (C.c_int, [C.POINTER(Dims), C.c_float, C.c_int, C.c_uint32, _fp, _fp, _fp, _fp, _fp, _fp,
                                 C.POINTER(Obs), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "visde_profile_begin": (C.c_int, [C.c_int]),
    "visde_profile_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "visde_session_create": (C.c_int, [C.POINTER(Dims), C.c_int, C.c_uint32, C.c_int32, C.c_int32, C.c_int32,
                                       C.POINTER(UserSde), C.POINTER(_fp)]),
    "visde_session_destroy": (None, [_fp]),
    "visde_session_h2d_bytes": (C.c_size_t, [_fp]),
    "visde_session_d2h_bytes": (C.c_size_t, [_fp]),
    "visde_session_launches": (C.c_int, [_fp]),
    "visde_session_set_noise_seed": (C.c_int, [_fp, C.c_uint64]),
    "visde_session_step": (C.c_int, [_fp, C.c_float, _fp, _fp, _fp, _fp, C.POINTER(Weights), C.POINTER(Obs), _fp, _fp,
                                     _fp, C.POINTER(Weights), _fp]),
    "visde_session_submit": (C.c_int, [_fp, C.c_float, _fp, _fp, _fp, _fp, C.POINTER(Weights), C.POINTER(Obs), _fp, _fp,
                                       _fp, C.POINTER(Weights), _fp]),
    "visde_session_wait": (C.c_int, [_fp]),
    "visde_em_fwd": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_uint32, C.c_float, _fp, _fp, _fp, C.c_uint64, _fp, _fp]),
    "visde_em_bwd": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_uint32, C.c_float, _fp, _fp, C.c_uint64, _fp, _fp,
                               _fp, _fp, _fp]),
    "visde_philox_normal": (C.c_int, [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, _fp, _fp]),
    "visde_path_summary_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int32]),
    "visde_path_summary": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_uint32, _fp, _fp, _fp, _fp, _fp, C.c_size_t,
                                     _fp]),
    "visde_grad_sqnorm_workspace_bytes": (C.c_size_t, []),
    "visde_grad_sqnorm": (C.c_int, [C.c_int64, _fp, _fp, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    "visde_adamw_ema_step": (C.c_int, [C.c_int64, _fp, _fp, _fp, _fp, _fp, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_int64, C.c_float, _fp, _fp, C.c_float, _fp, _fp]),
}

_lib: C.CDLL | None = None


def build(verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", str(PKG / "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libvisde.so failed")
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the path-sampling kernels)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.visde_version() != VERSION:
            raise RuntimeError("libvisde.so version mismatch; rebuild")
        _lib = lib
    return _lib


def check(rc: int) -> None:
    """0 -> ok; EINVAL -> ValueError (reference raises ValueError on bad config,
    models/head.py:33-36, kernels/weights.py:89-90); anything else -> RuntimeError."""
    if rc == OK:
        return
    msg = load().visde_last_error().decode()
    if rc == EINVAL:
        raise ValueError(msg)
    raise RuntimeError(f"libvisde error {rc}: {msg}")
