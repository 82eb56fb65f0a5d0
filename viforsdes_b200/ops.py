"""PyTorch operators over the libvisde C ABI.

``torch.library`` custom ops (namespace ``visde``) replace the reference's autograd glue
``_SDEFunction`` (src/variational_sde/kernels/autograd.py:35-241) and its launchers
``launch_fwd`` / ``launch_bwd`` (kernels/forward.py:378, kernels/backward.py:627):

    visde::path_fwd   -> paths, means, chol, stash        (K0 context GEMM + K1 recurrence)
    visde::path_bwd   -> grad_x0, grad_context, grad_theta, 4*NL+2 weight grads (K2 + K3 + K4)
    visde::elbo_fwd / visde::elbo_bwd                      (K5 / K6 time-parallel ELBO terms)

PyTorch is used for device memory, streams and autograd plumbing only; all math runs in the
CUDA kernels of ``csrc/``.  There is no CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from viforsdes_b200 import _lib

_VARIANT = _lib.VARIANT_AUTO  # tests flip this to cross-check the two kernel families


def set_variant(v: int) -> None:
    global _VARIANT
    _VARIANT = int(v)


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "viforsdes_b200 ops run on CUDA tensors only (no CPU fallback; the reference has none "
                "either: models/head.py:164-209 always launches its GPU kernels)")


def _f32c(t: Tensor) -> Tensor:
    if t.dtype == torch.float32 and not t.requires_grad and t.is_contiguous():
        return t  # the common case (fp32 parameters / buffers handed over by autograd): no dispatcher round trips
    return t.detach().to(torch.float32).contiguous()


_SIZES: dict = {}      # (dims..., variant, backward) -> bytes: the size queries are pure functions of the dims
_WORKSPACE: dict = {}  # (device index, stream) -> uint8 scratch, grown on demand, reused by every call on that stream


def _stash_bytes(lib, d: _lib.Dims) -> int:
    key = (d.B, d.T, d.S, d.C, d.P, d.H, d.NL, d.variant, "stash")
    if key not in _SIZES:
        _SIZES[key] = int(lib.visde_stash_bytes(C.byref(d)))
    return _SIZES[key]


def _workspace(lib, d: _lib.Dims, backward: int, dev: torch.device) -> Tuple[Tensor, int]:
    """Scratch for one launch.  Kernels on one stream run in order and the scratch holds nothing across calls, so one
    buffer per (device, stream) serves every call (the reference allocates ~15 tensors per launch, kernels/forward.py:
    398-497); under CUDA-graph capture the allocation goes to the graph's private pool instead."""
    key = (d.B, d.T, d.S, d.C, d.P, d.H, d.NL, d.variant, backward)
    if key not in _SIZES:
        _SIZES[key] = int(lib.visde_workspace_bytes(C.byref(d), backward))
    need = _SIZES[key]
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(need, device=dev, dtype=torch.uint8), need
    slot = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream())
    buf = _WORKSPACE.get(slot)
    if buf is None or buf.numel() < need:
        buf = _WORKSPACE[slot] = torch.empty(need, device=dev, dtype=torch.uint8)
    return buf, need


def _dims(B: int, T: int, S: int, Cdim: int, P: int, H: int, NL: int) -> _lib.Dims:
    return _lib.Dims(B, T, S, Cdim, P, H, NL, _VARIANT)


def _weights_struct(w_ih: Sequence[Tensor], w_hh: Sequence[Tensor], b_ih: Sequence[Tensor], b_hh: Sequence[Tensor],
                    out_w: Tensor, out_b: Tensor) -> _lib.Weights:
    w = _lib.Weights()
    for k in range(len(w_hh)):
        w.w_ih[k] = w_ih[k].data_ptr()
        w.w_hh[k] = w_hh[k].data_ptr()
        w.b_ih[k] = b_ih[k].data_ptr()
        w.b_hh[k] = b_hh[k].data_ptr()
    w.out_w = out_w.data_ptr()
    w.out_b = out_b.data_ptr()
    return w


def _ctx_view(ctx: Tensor) -> Tuple[Tensor, _lib.CtxView]:
    """Strided [B,T,C] view consumed in place (fp32 or bf16, innermost stride 1)."""
    if ctx.dtype not in (torch.float32, torch.bfloat16):
        ctx = ctx.to(torch.float32)
    if ctx.dim() != 3:
        raise ValueError(f"context must be [B,T,C], got {tuple(ctx.shape)}")
    if ctx.shape[2] > 0 and ctx.stride(2) != 1:
        ctx = ctx.contiguous()
    dt = _lib.BF16 if ctx.dtype == torch.bfloat16 else _lib.F32
    return ctx, _lib.CtxView(_ptr(ctx), ctx.stride(0), ctx.stride(1), dt)


def _shapes(x0: Tensor, context: Tensor, theta: Tensor, eps: Tensor, w_ih, w_hh, out_w):
    B, S = x0.shape
    T, Cdim = context.shape[1], context.shape[2]
    P = theta.shape[1]
    NL = len(w_hh)
    if NL < 1 or NL > _lib.MAX_LAYERS:
        raise ValueError(f"num_layers must be in [1, {_lib.MAX_LAYERS}], got {NL}")  # models/head.py:33-36
    H = w_hh[0].shape[1]
    if context.shape[0] != B or theta.shape[0] != B or tuple(eps.shape) != (B, T, S):
        raise ValueError("inconsistent batch / step / state dims between x0, context, theta, noise")
    if tuple(w_ih[0].shape) != (3 * H, S + Cdim + P):
        raise ValueError(f"weight_ih_l0 must be [{3 * H}, {S + Cdim + P}], got {tuple(w_ih[0].shape)}")
    if tuple(out_w.shape) != (S + S * (S + 1) // 2, H):
        raise ValueError(f"out_proj.weight must be [{S + S * (S + 1) // 2}, {H}], got {tuple(out_w.shape)}")
    return B, T, S, Cdim, P, H, NL


# ------------------------------------------------------------------------------------------
# visde::path_fwd / visde::path_bwd
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("visde::path_fwd", mutates_args=())
def path_fwd(x0: Tensor, context: Tensor, theta: Tensor, eps: Tensor, w_ih: List[Tensor], w_hh: List[Tensor],
             b_ih: List[Tensor], b_hh: List[Tensor], out_w: Tensor, out_b: Tensor, dt: float,
             save: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    _require_cuda(x0, context, theta, eps, out_w)
    B, T, S, Cdim, P, H, NL = _shapes(x0, context, theta, eps, w_ih, w_hh, out_w)
    lib = _lib.load()
    dev = x0.device
    x0f, thf, epf = _f32c(x0), _f32c(theta), _f32c(eps)
    ctx, cv = _ctx_view(context.detach())
    ws_ih, ws_hh = [_f32c(t) for t in w_ih], [_f32c(t) for t in w_hh]
    bs_ih, bs_hh = [_f32c(t) for t in b_ih], [_f32c(t) for t in b_hh]
    ow, ob = _f32c(out_w), _f32c(out_b)
    w = _weights_struct(ws_ih, ws_hh, bs_ih, bs_hh, ow, ob)
    d = _dims(B, T, S, Cdim, P, H, NL)
    paths = torch.empty(B, T + 1, S, device=dev, dtype=torch.float32)
    means = torch.empty(B, T, S, device=dev, dtype=torch.float32)
    chol = torch.empty(B, T, S, S, device=dev, dtype=torch.float32)
    stash = torch.empty(_stash_bytes(lib, d) if save else 0, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        ws, ws_bytes = _workspace(lib, d, 0, dev)
        _lib.check(lib.visde_path_fwd(C.byref(d), dt, _ptr(x0f), C.byref(cv), _ptr(thf), _ptr(epf), C.byref(w),
                                      _ptr(paths), _ptr(means), _ptr(chol), _ptr(stash) if save else None,
                                      _ptr(ws), ws_bytes, _stream()))
    return paths, means, chol, stash


@path_fwd.register_fake
def _(x0, context, theta, eps, w_ih, w_hh, b_ih, b_hh, out_w, out_b, dt, save):
    B, S = x0.shape
    T = context.shape[1]
    f = dict(device=x0.device, dtype=torch.float32)
    return (torch.empty(B, T + 1, S, **f), torch.empty(B, T, S, **f), torch.empty(B, T, S, S, **f),
            torch.empty(0, device=x0.device, dtype=torch.uint8))


@torch.library.custom_op("visde::path_bwd", mutates_args=())
def path_bwd(g_paths: Tensor, g_means: Tensor, g_chol: Tensor, context: Tensor, theta: Tensor, eps: Tensor,
             paths: Tensor, stash: Tensor, w_ih: List[Tensor], w_hh: List[Tensor], b_ih: List[Tensor],
             b_hh: List[Tensor], out_w: Tensor, out_b: Tensor, dt: float) -> List[Tensor]:
    """Returns [grad_x0, grad_context, grad_theta, gw_ih*NL, gw_hh*NL, gb_ih*NL, gb_hh*NL, g_out_w, g_out_b]."""
    _require_cuda(g_paths, context, theta, eps, paths, stash)
    B, T, S = eps.shape
    Cdim, P = context.shape[2], theta.shape[1]
    NL, H = len(w_hh), w_hh[0].shape[1]
    lib = _lib.load()
    dev = eps.device
    thf, epf = _f32c(theta), _f32c(eps)
    ctx, cv = _ctx_view(context.detach())
    ws_ih, ws_hh = [_f32c(t) for t in w_ih], [_f32c(t) for t in w_hh]
    bs_ih, bs_hh = [_f32c(t) for t in b_ih], [_f32c(t) for t in b_hh]
    ow, ob = _f32c(out_w), _f32c(out_b)
    w = _weights_struct(ws_ih, ws_hh, bs_ih, bs_hh, ow, ob)
    d = _dims(B, T, S, Cdim, P, H, NL)
    f = dict(device=dev, dtype=torch.float32)
    grad_x0 = torch.empty(B, S, **f)
    grad_ctx = torch.empty(B, T, Cdim, device=dev, dtype=ctx.dtype)
    grad_theta = torch.empty(B, P, **f)
    gw_ih = [torch.empty_like(t) for t in ws_ih]
    gw_hh = [torch.empty_like(t) for t in ws_hh]
    gb_ih = [torch.empty_like(t) for t in bs_ih]
    gb_hh = [torch.empty_like(t) for t in bs_hh]
    g_ow, g_ob = torch.empty_like(ow), torch.empty_like(ob)
    gw = _weights_struct(gw_ih, gw_hh, gb_ih, gb_hh, g_ow, g_ob)
    gv = _lib.CtxView(_ptr(grad_ctx), grad_ctx.stride(0), grad_ctx.stride(1), cv.dtype)
    with torch.cuda.device(dev):
        ws, ws_bytes = _workspace(lib, d, 1, dev)
        _lib.check(lib.visde_path_bwd(C.byref(d), dt, _ptr(_f32c(g_paths)), _ptr(_f32c(g_means)), _ptr(_f32c(g_chol)),
                                      C.byref(cv), _ptr(thf), _ptr(epf), C.byref(w), _ptr(_f32c(paths)), _ptr(stash),
                                      _ptr(grad_x0), C.byref(gv), _ptr(grad_theta), C.byref(gw), _ptr(ws), ws_bytes,
                                      _stream()))
    return [grad_x0, grad_ctx, grad_theta, *gw_ih, *gw_hh, *gb_ih, *gb_hh, g_ow, g_ob]


@path_bwd.register_fake
def _(g_paths, g_means, g_chol, context, theta, eps, paths, stash, w_ih, w_hh, b_ih, b_hh, out_w, out_b, dt):
    B, T, S = eps.shape
    f = dict(device=eps.device, dtype=torch.float32)
    return [torch.empty(B, S, **f), torch.empty(B, T, context.shape[2], device=eps.device, dtype=context.dtype),
            torch.empty(B, theta.shape[1], **f), *[torch.empty_like(t) for t in w_ih],
            *[torch.empty_like(t) for t in w_hh], *[torch.empty_like(t) for t in b_ih],
            *[torch.empty_like(t) for t in b_hh], torch.empty_like(out_w), torch.empty_like(out_b)]


def _path_setup(ctx, inputs, output):
    x0, context, theta, eps, w_ih, w_hh, b_ih, b_hh, out_w, out_b, dt, save = inputs
    paths, _means, _chol, stash = output
    if not save:
        return
    ctx.save_for_backward(context, theta, eps, paths, stash, *w_ih, *w_hh, *b_ih, *b_hh, out_w, out_b)
    ctx.nl = len(w_hh)
    ctx.dt = dt
    ctx.dtypes = (x0.dtype, context.dtype, theta.dtype)


def _path_backward(ctx, g_paths, g_means, g_chol, _g_stash):
    if not hasattr(ctx, "nl"):
        raise RuntimeError("visde::path_fwd was called with save=False; no activations to differentiate")
    nl = ctx.nl
    context, theta, eps, paths, stash, *rest = ctx.saved_tensors
    w_ih, w_hh, b_ih, b_hh = rest[:nl], rest[nl:2 * nl], rest[2 * nl:3 * nl], rest[3 * nl:4 * nl]
    out_w, out_b = rest[4 * nl], rest[4 * nl + 1]
    z = lambda like: torch.zeros_like(like, dtype=torch.float32)  # noqa: E731
    g_paths = z(paths) if g_paths is None else g_paths
    B, T, S = eps.shape
    g_means = torch.zeros(B, T, S, device=eps.device) if g_means is None else g_means
    g_chol = torch.zeros(B, T, S, S, device=eps.device) if g_chol is None else g_chol
    out = torch.ops.visde.path_bwd(g_paths, g_means, g_chol, context, theta, eps, paths, stash, list(w_ih),
                                   list(w_hh), list(b_ih), list(b_hh), out_w, out_b, ctx.dt)
    x0_dt, ctx_dt, th_dt = ctx.dtypes
    gx0, gctx, gth = out[0].to(x0_dt), out[1].to(ctx_dt), out[2].to(th_dt)
    o = out[3:]
    cast = lambda gs, ws: [g.to(w.dtype) for g, w in zip(gs, ws)]  # noqa: E731
    return (gx0, gctx, gth, None, cast(o[:nl], w_ih), cast(o[nl:2 * nl], w_hh), cast(o[2 * nl:3 * nl], b_ih),
            cast(o[3 * nl:4 * nl], b_hh), o[4 * nl].to(out_w.dtype), o[4 * nl + 1].to(out_b.dtype), None, None)


path_fwd.register_autograd(_path_backward, setup_context=_path_setup)


# ------------------------------------------------------------------------------------------
# visde::elbo_fwd / visde::elbo_bwd
# ------------------------------------------------------------------------------------------
def _obs_struct(obs_idx: Tensor, obs_values: Tensor, obs_matrix: Optional[Tensor], variance: float):
    idx = obs_idx.detach().to(torch.int32).contiguous()
    vals = _f32c(obs_values)
    mat = None if obs_matrix is None else _f32c(obs_matrix)
    n_obs = idx.shape[0]
    obs_dim = vals.shape[1] if vals.dim() == 2 else 0
    o = _lib.Obs(n_obs, obs_dim, _ptr(idx), _ptr(vals), _ptr(mat), float(variance))
    return o, (idx, vals, mat)


@torch.library.custom_op("visde::elbo_fwd", mutates_args=())
def elbo_fwd(z: Tensor, means: Tensor, chol: Tensor, theta: Tensor, drift: Optional[Tensor],
             diffusion: Optional[Tensor], obs_idx: Tensor, obs_values: Tensor, obs_matrix: Optional[Tensor],
             obs_variance: float, dt: float, sde_kind: int, positive_mask: int) -> Tensor:
    """terms [B,4] = per-trajectory (obs, sde, gen, jacobian) log-probabilities."""
    _require_cuda(z, means, chol, theta)
    lib = _lib.load()
    B, T1, S = z.shape
    T = T1 - 1
    P = theta.shape[1]
    d = _dims(B, T, S, 0, P, 1, 1)
    zf, mf, cf, tf = _f32c(z), _f32c(means), _f32c(chol), _f32c(theta)
    df = None if drift is None else _f32c(drift)
    Df = None if diffusion is None else _f32c(diffusion)
    o, keep = _obs_struct(obs_idx, obs_values, obs_matrix, obs_variance)
    terms = torch.empty(B, 4, device=z.device, dtype=torch.float32)
    with torch.cuda.device(z.device):
        _lib.check(lib.visde_elbo_fwd(C.byref(d), dt, sde_kind, positive_mask, _ptr(zf), _ptr(mf), _ptr(cf), _ptr(tf),
                                      _ptr(df), _ptr(Df), C.byref(o), _ptr(terms), _stream()))
    del keep
    return terms


@elbo_fwd.register_fake
def _(z, means, chol, theta, drift, diffusion, obs_idx, obs_values, obs_matrix, obs_variance, dt, sde_kind,
      positive_mask):
    return torch.empty(z.shape[0], 4, device=z.device, dtype=torch.float32)


@torch.library.custom_op("visde::elbo_bwd", mutates_args=())
def elbo_bwd(g_terms: Tensor, z: Tensor, means: Tensor, chol: Tensor, theta: Tensor, drift: Optional[Tensor],
             diffusion: Optional[Tensor], obs_idx: Tensor, obs_values: Tensor, obs_matrix: Optional[Tensor],
             obs_variance: float, dt: float, sde_kind: int, positive_mask: int) -> List[Tensor]:
    """Returns [g_z, g_means, g_chol, g_theta, g_drift, g_diffusion] (last two empty unless generic)."""
    _require_cuda(g_terms, z, means, chol, theta)
    lib = _lib.load()
    B, T1, S = z.shape
    T = T1 - 1
    P = theta.shape[1]
    d = _dims(B, T, S, 0, P, 1, 1)
    zf, mf, cf, tf = _f32c(z), _f32c(means), _f32c(chol), _f32c(theta)
    df = None if drift is None else _f32c(drift)
    Df = None if diffusion is None else _f32c(diffusion)
    o, keep = _obs_struct(obs_idx, obs_values, obs_matrix, obs_variance)
    f = dict(device=z.device, dtype=torch.float32)
    g_z, g_m, g_c = torch.empty(B, T + 1, S, **f), torch.empty(B, T, S, **f), torch.empty(B, T, S, S, **f)
    g_th = torch.empty(B, P, **f)
    generic = sde_kind == _lib.SDE_GENERIC
    g_dr = torch.empty(B, T, S, **f) if generic else torch.empty(0, **f)
    g_di = torch.empty(B, T, S, S, **f) if generic else torch.empty(0, **f)
    with torch.cuda.device(z.device):
        _lib.check(lib.visde_elbo_bwd(C.byref(d), dt, sde_kind, positive_mask, _ptr(zf), _ptr(mf), _ptr(cf), _ptr(tf),
                                      _ptr(df), _ptr(Df), C.byref(o), _ptr(_f32c(g_terms)), _ptr(g_z), _ptr(g_m),
                                      _ptr(g_c), _ptr(g_th), _ptr(g_dr), _ptr(g_di), _stream()))
    del keep
    return [g_z, g_m, g_c, g_th, g_dr, g_di]


@elbo_bwd.register_fake
def _(g_terms, z, means, chol, theta, drift, diffusion, obs_idx, obs_values, obs_matrix, obs_variance, dt, sde_kind,
      positive_mask):
    B, T1, S = z.shape
    f = dict(device=z.device, dtype=torch.float32)
    generic = sde_kind == _lib.SDE_GENERIC
    return [torch.empty(B, T1, S, **f), torch.empty(B, T1 - 1, S, **f), torch.empty(B, T1 - 1, S, S, **f),
            torch.empty(B, theta.shape[1], **f), torch.empty(B, T1 - 1, S, **f) if generic else torch.empty(0, **f),
            torch.empty(B, T1 - 1, S, S, **f) if generic else torch.empty(0, **f)]


def _elbo_setup(ctx, inputs, output):
    (z, means, chol, theta, drift, diffusion, obs_idx, obs_values, obs_matrix, obs_variance, dt, sde_kind,
     positive_mask) = inputs
    ctx.save_for_backward(z, means, chol, theta, drift, diffusion, obs_idx, obs_values, obs_matrix)
    ctx.consts = (obs_variance, dt, sde_kind, positive_mask)
    ctx.dtypes = (z.dtype, means.dtype, chol.dtype, theta.dtype)


def _elbo_backward(ctx, g_terms):
    z, means, chol, theta, drift, diffusion, obs_idx, obs_values, obs_matrix = ctx.saved_tensors
    obs_variance, dt, sde_kind, positive_mask = ctx.consts
    g = torch.ops.visde.elbo_bwd(g_terms.contiguous(), z, means, chol, theta, drift, diffusion, obs_idx, obs_values,
                                 obs_matrix, obs_variance, dt, sde_kind, positive_mask)
    generic = sde_kind == _lib.SDE_GENERIC
    zd, md, cd, td = ctx.dtypes
    return (g[0].to(zd), g[1].to(md), g[2].to(cd), g[3].to(td), g[4].to(drift.dtype) if generic else None,
            g[5].to(diffusion.dtype) if generic else None, None, None, None, None, None, None, None)


elbo_fwd.register_autograd(_elbo_backward, setup_context=_elbo_setup)
