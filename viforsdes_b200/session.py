"""Host-buffer iteration through the C ABI (``visde_session_*`` in include/visde.h): the call a
non-PyTorch host (or the reference's trainer, inference/trainer.py:176-198, via ctypes) makes for
one ELBO iteration of the path: pinned host tensors in, ELBO terms and gradients out."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
from torch import Tensor

from viforsdes_b200 import _lib


def _pin(t: Tensor, dtype: torch.dtype = torch.float32) -> Tensor:
    t = t.detach().to(dtype).contiguous().cpu()
    return t.pin_memory() if torch.cuda.is_available() else t


class _DeviceArray:
    """A raw device pointer handed to a hook, exposed through ``__cuda_array_interface__`` so that
    ``torch.as_tensor`` aliases it (no copy)."""

    def __init__(self, ptr: int, shape) -> None:
        self.__cuda_array_interface__ = {"shape": tuple(int(n) for n in shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _alias(ptr: int, shape, device) -> Tensor:
    return torch.as_tensor(_DeviceArray(ptr, shape), device=device)


class _UserSdeHooks:
    """ctypes trampolines of ``visde_user_sde``: run the user's PyTorch ``drift`` / ``diffusion`` (SDE protocol,
    src/variational_sde/core/sde.py:8-15) and their autograd vector-Jacobian product on the session's device
    tensors and stream -- the reference's evidence_lower_bound.py:37-40 call, moved behind the C ABI.

    The session's device buffers never move, so after `warmup` eager iterations the two hook bodies are captured as a pair of
    CUDA graphs sharing one memory pool (forward graph keeps the autograd graph's tensors alive for the backward graph, the
    pattern of ``torch.cuda.make_graphed_callables``) and replayed: the ~35 small PyTorch kernels + autograd bookkeeping of a
    Lorenz-96 iteration cost ~2 ms of CPU time per iteration eagerly, which the GPU waits for.  A user SDE whose code cannot be
    captured (host synchronisation, data-dependent control flow) stays eager."""

    def __init__(self, sde, B: int, T: int, S: int, P: int, graphs: bool = True, warmup: int = 3) -> None:
        self.sde, self.B, self.T, self.S, self.P = sde, B, T, S, P
        self.error: BaseException | None = None
        self._drift = self._diffusion = self._x = self._th = None
        self.struct = _lib.UserSde(_lib.SDE_EVAL_FN(self._eval), _lib.SDE_VJP_FN(self._vjp), None)
        self.want_graphs, self.warmup, self.calls = graphs, warmup, 0
        self.g_eval = self.g_vjp = None
        self._ptrs = None  # device addresses the graphs were captured for
        self._theta_static = None

    # -- bodies (run eagerly or under capture) -----------------------------------------------------------------------
    def _eval_body(self, x_ptr, drift_ptr, diff_ptr, dev) -> None:
        B, T, S, P = self.B, self.T, self.S, self.P
        with torch.enable_grad():
            self._x = _alias(x_ptr, (B * T, S), dev).requires_grad_(True)
            self._th = self._theta_static.detach().requires_grad_(True)
            th_flat = self._th[:, None, :].expand(B, T, P).reshape(B * T, P)
            self._drift = self.sde.drift(self._x, th_flat)
            self._diffusion = self.sde.diffusion(self._x, th_flat)
        with torch.no_grad():
            _alias(drift_ptr, (B * T, S), dev).copy_(self._drift.reshape(B * T, S))
            _alias(diff_ptr, (B * T, S, S), dev).copy_(self._diffusion.reshape(B * T, S, S))

    def _vjp_body(self, g_drift_ptr, g_diff_ptr, g_x_ptr, g_th_ptr, dev) -> None:
        B, T, S, P = self.B, self.T, self.S, self.P
        outs, gouts = [], []
        for o, ptr in ((self._drift, g_drift_ptr), (self._diffusion, g_diff_ptr)):
            if o.requires_grad:
                outs.append(o)
                gouts.append(_alias(ptr, tuple(o.shape), dev).to(o.dtype))
        gx = gth = None
        if outs:
            gx, gth = torch.autograd.grad(outs, [self._x, self._th], gouts, allow_unused=True)
        gx_out, gth_out = _alias(g_x_ptr, (B * T, S), dev), _alias(g_th_ptr, (B, P), dev)
        gx_out.zero_() if gx is None else gx_out.copy_(gx)
        gth_out.zero_() if gth is None else gth_out.copy_(gth)

    # -- trampolines --------------------------------------------------------------------------------------------------
    def _eval(self, _user, x_ptr, th_ptr, drift_ptr, diff_ptr, stream) -> int:
        try:
            dev = torch.device("cuda", torch.cuda.current_device())
            ext = torch.cuda.ExternalStream(stream or 0)
            with torch.cuda.stream(ext):
                if self._theta_static is None:
                    self._theta_static = torch.empty(self.B, self.P, device=dev)
                # theta lives in one of the session's two input sets: stage it at a fixed address for the graphs
                self._theta_static.copy_(_alias(th_ptr, (self.B, self.P), dev))
                self.calls += 1
                ptrs = (x_ptr, drift_ptr, diff_ptr)
                if self.g_eval is not None and self._ptrs[:3] == ptrs:
                    self.g_eval.replay()
                    return 0
                self._capturing = False
                if self.want_graphs and self.g_eval is None and self.calls > self.warmup:
                    try:
                        torch.cuda.synchronize()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=ext):
                            self._eval_body(x_ptr, drift_ptr, diff_ptr, dev)
                        self.g_eval, self._ptrs, self._capturing = g, ptrs, True
                        g.replay()
                        return 0
                    except Exception:  # not capturable: stay eager from now on
                        self.want_graphs, self.g_eval = False, None
                        torch.cuda.synchronize()
                self._eval_body(x_ptr, drift_ptr, diff_ptr, dev)
            return 0
        except BaseException as e:  # noqa: BLE001  (must not unwind through the C frame)
            self.error = e
            return -1

    def _vjp(self, _user, x_ptr, th_ptr, g_drift_ptr, g_diff_ptr, g_x_ptr, g_th_ptr, stream) -> int:
        try:
            dev = torch.device("cuda", torch.cuda.current_device())
            ext = torch.cuda.ExternalStream(stream or 0)
            with torch.cuda.stream(ext):
                ptrs = (g_drift_ptr, g_diff_ptr, g_x_ptr, g_th_ptr)
                if self.g_vjp is not None and self._ptrs[3:] == ptrs:
                    self.g_vjp.replay()
                    return 0
                if self.g_eval is not None and getattr(self, "_capturing", False):
                    try:
                        torch.cuda.synchronize()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=ext, pool=self.g_eval.pool()):
                            self._vjp_body(g_drift_ptr, g_diff_ptr, g_x_ptr, g_th_ptr, dev)
                        self.g_vjp, self._ptrs, self._capturing = g, self._ptrs[:3] + ptrs, False
                        g.replay()
                        return 0
                    except Exception:
                        self.want_graphs, self.g_eval, self.g_vjp = False, None, None
                        torch.cuda.synchronize()
                        self._eval_body(*self._ptrs[:3], dev)  # rebuild the autograd graph eagerly for this iteration
                self._vjp_body(g_drift_ptr, g_diff_ptr, g_x_ptr, g_th_ptr, dev)
                if self.g_eval is None:
                    self._drift = self._diffusion = self._x = self._th = None
            return 0
        except BaseException as e:  # noqa: BLE001
            self.error = e
            return -1


class HostSession:
    def __init__(self, *, x0: Tensor, context_full: Tensor, theta: Tensor, eps: Tensor, w_ih: List[Tensor],
                 w_hh: List[Tensor], b_ih: List[Tensor], b_hh: List[Tensor], out_w: Tensor, out_b: Tensor, dt: float,
                 sde_kind: int, positive_mask: int, obs_idx: Tensor, obs_values: Tensor, obs_variance: float,
                 variant: int = _lib.VARIANT_AUTO, want_grad_context: bool = False, sde=None,
                 context_dtype: torch.dtype = torch.float32, graph_hooks: bool = True,
                 device_noise_seed: Optional[int] = None) -> None:
        """`context_dtype` bfloat16 = the reference's AMP mode (the encoder runs under bf16 autocast,
        inference/trainer.py:171-175): the host context and grad_context are bf16, which halves the dominant H2D term.
        `sde`: the user SDE object when `sde_kind` is GENERIC.  `device_noise_seed`: draw the noise on the device as the
        reference's sampler does (inference/diffusion_path_sampler.py:57) -- iteration i uses the Philox stream of
        ``visde_philox_normal(seed + i)`` and `eps` never crosses the bus (it is then only a shape)."""
        self.lib = _lib.load()
        if context_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("context_dtype must be float32 or bfloat16")
        B, S = x0.shape
        T, Cd = context_full.shape[1] - 1, context_full.shape[2]
        P, NL, H = theta.shape[1], len(w_hh), w_hh[0].shape[1]
        self.dims = _lib.Dims(B, T, S, Cd, P, H, NL, variant)
        self.dt = float(dt)
        self.context_dtype = context_dtype
        self.x0, self.ctx, self.theta, self.eps = _pin(x0), _pin(context_full, context_dtype), _pin(theta), _pin(eps)
        self.w = [[_pin(t) for t in ws] for ws in (w_ih, w_hh, b_ih, b_hh)]
        self.out_w, self.out_b = _pin(out_w), _pin(out_b)
        # two host output sets: iteration i+1 may be submitted before the outputs of i are consumed
        self._out = [self._make_outputs(B, S, P, T, Cd, want_grad_context) for _ in range(2)]
        self._submitted = self._waited = 0
        self.device_noise_seed = device_noise_seed
        self.obs_idx = obs_idx.detach().to(torch.int32).contiguous().cpu()
        self.obs_values = obs_values.detach().to(torch.float32).contiguous().cpu()
        self.obs = _lib.Obs(self.obs_idx.shape[0], self.obs_values.shape[1], self.obs_idx.data_ptr(),
                            self.obs_values.data_ptr(), None, float(obs_variance))
        self.handle = C.c_void_p()
        self._hooks = None
        if sde_kind == _lib.SDE_GENERIC:
            if sde is None:
                raise ValueError("sde_kind GENERIC needs the user SDE object (sde=...)")
            self._hooks = _UserSdeHooks(sde, B, T, S, P, graphs=graph_hooks)
        _lib.check(self.lib.visde_session_create(
            C.byref(self.dims), sde_kind, positive_mask, self.obs.n_obs, self.obs.obs_dim,
            _lib.BF16 if context_dtype == torch.bfloat16 else _lib.F32,
            C.byref(self._hooks.struct) if self._hooks else None, C.byref(self.handle)))
        if device_noise_seed is not None:
            _lib.check(self.lib.visde_session_set_noise_seed(self.handle, int(device_noise_seed)))

    @classmethod
    def from_problem(cls, p, **kw) -> "HostSession":
        """Build from an oracle ``Problem`` (tests / bench only construct the inputs there)."""
        B, T, Cd = p.context.shape
        full = torch.zeros(B, T + 1, Cd)
        full[:, :T] = p.context
        w = p.weights
        kind = {"ou": _lib.SDE_OU, "lv": _lib.SDE_LV}.get(p.name, _lib.SDE_GENERIC)
        if kind == _lib.SDE_GENERIC:
            kw.setdefault("sde", p.sde)
        mask = 0
        for d in p.positive_dims:
            mask |= 1 << d
        idx = torch.clamp(torch.round(p.obs_times / p.dt).long(), max=T)
        return cls(x0=p.x0, context_full=full, theta=p.theta, eps=p.eps, w_ih=w.w_ih, w_hh=w.w_hh, b_ih=w.b_ih,
                   b_hh=w.b_hh, out_w=w.out_w, out_b=w.out_b, dt=p.dt, sde_kind=kind, positive_mask=mask,
                   obs_idx=idx, obs_values=p.obs_values, obs_variance=p.obs_variance, **kw)

    @classmethod
    def from_inputs(cls, inp, **kw) -> "HostSession":
        """Build from ``viforsdes_b200.synthetic.Inputs``."""
        return cls(x0=inp.x0, context_full=inp.context_full, theta=inp.theta, eps=inp.eps, w_ih=inp.w_ih,
                   w_hh=inp.w_hh, b_ih=inp.b_ih, b_hh=inp.b_hh, out_w=inp.out_w, out_b=inp.out_b, dt=inp.dt,
                   sde_kind=inp.sde_kind, positive_mask=inp.positive_mask, obs_idx=inp.obs_idx,
                   obs_values=inp.obs_values, obs_variance=inp.obs_variance, sde=inp.sde, **kw)

    def _wstruct(self, groups, ow, ob) -> _lib.Weights:
        s = _lib.Weights()
        for k in range(self.dims.NL):
            s.w_ih[k], s.w_hh[k] = groups[0][k].data_ptr(), groups[1][k].data_ptr()
            s.b_ih[k], s.b_hh[k] = groups[2][k].data_ptr(), groups[3][k].data_ptr()
        s.out_w, s.out_b = ow.data_ptr(), ob.data_ptr()
        return s

    @property
    def h2d_bytes(self) -> int:
        return int(self.lib.visde_session_h2d_bytes(self.handle))

    @property
    def d2h_bytes(self) -> int:
        g = self._out[0]["grad_ctx"]
        extra = g.numel() * g.element_size() if g is not None else 0
        return int(self.lib.visde_session_d2h_bytes(self.handle)) + extra

    @property
    def launches(self) -> int:
        return int(self.lib.visde_session_launches(self.handle))

    def _make_outputs(self, B: int, S: int, P: int, T: int, Cd: int, want_grad_context: bool) -> Dict[str, object]:
        pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
        return {
            "gw": [[pin(torch.empty_like(t)) for t in ws] for ws in self.w],
            "g_out_w": pin(torch.empty_like(self.out_w)), "g_out_b": pin(torch.empty_like(self.out_b)),
            "terms": pin(torch.empty(B, 4)), "grad_x0": pin(torch.empty(B, S)), "grad_theta": pin(torch.empty(B, P)),
            "grad_ctx": pin(torch.empty(B, T + 1, Cd, dtype=self.context_dtype)) if want_grad_context else None,
        }

    def _call(self, fn, o) -> None:
        w = self._wstruct(self.w, self.out_w, self.out_b)
        gw = self._wstruct(o["gw"], o["g_out_w"], o["g_out_b"])
        rc = fn(
            self.handle, self.dt, self.x0.data_ptr(), self.ctx.data_ptr(), self.theta.data_ptr(),
            None if self.device_noise_seed is not None else self.eps.data_ptr(),
            C.byref(w), C.byref(self.obs), o["terms"].data_ptr(), o["grad_x0"].data_ptr(), o["grad_theta"].data_ptr(),
            C.byref(gw), None if o["grad_ctx"] is None else o["grad_ctx"].data_ptr())
        if rc and self._hooks is not None and self._hooks.error is not None:
            err, self._hooks.error = self._hooks.error, None
            raise err  # the user's drift / diffusion raised inside a hook
        _lib.check(rc)

    def _results(self, o) -> Dict[str, object]:
        grads = {"x0": o["grad_x0"], "theta": o["grad_theta"], "out_w": o["g_out_w"], "out_b": o["g_out_b"]}
        for k in range(self.dims.NL):
            grads[f"w_ih_l{k}"], grads[f"w_hh_l{k}"] = o["gw"][0][k], o["gw"][1][k]
            grads[f"b_ih_l{k}"], grads[f"b_hh_l{k}"] = o["gw"][2][k], o["gw"][3][k]
        if o["grad_ctx"] is not None:
            grads["context"] = o["grad_ctx"][:, : self.dims.T]
        return {"terms": o["terms"], "grads": grads}

    def step(self) -> Dict[str, object]:
        """One synchronous iteration (visde_session_step): H2D, kernels, D2H, wait."""
        o = self._out[self._submitted & 1]
        self._call(self.lib.visde_session_step, o)
        self._submitted += 1
        self._waited += 1
        return self._results(o)

    def submit(self) -> None:
        """Enqueue one iteration without waiting (visde_session_submit); at most two in flight.  The host
        inputs are read when the copies run: change them only after the matching wait()."""
        self._call(self.lib.visde_session_submit, self._out[self._submitted & 1])
        self._submitted += 1

    def wait(self) -> Dict[str, object]:
        """Block until the oldest in-flight iteration's outputs are in host memory and return them."""
        _lib.check(self.lib.visde_session_wait(self.handle))
        o = self._out[self._waited & 1]
        self._waited += 1
        return self._results(o)

    def close(self) -> None:
        if self.handle:
            self.lib.visde_session_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
