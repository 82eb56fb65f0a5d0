"""Device-resident ELBO iteration of the hot path, straight over the C ABI with buffers allocated
once (what a training loop does after its first step): K0 -> K1 -> K5 -> K6 -> K2 -> K3 -> K4.
Used by bench.py (`value` leg: inputs already in HBM) and by the data-parallel wrapper."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch
from torch import Tensor

from viforsdes_b200 import _lib
from viforsdes_b200.dist import FlatBucket
from viforsdes_b200.synthetic import Inputs


class PathIteration:
    def __init__(self, inp: Inputs, device: torch.device | str = "cuda", variant: int = _lib.VARIANT_AUTO,
                 context_dtype: torch.dtype = torch.float32) -> None:
        # user SDEs (sde_kind GENERIC): drift / diffusion and their vector-Jacobian products are evaluated in PyTorch on the
        # flattened [B*T, S] path between the kernels, as inference/evidence_lower_bound.py:37-40 does
        self.generic = inp.sde_kind == _lib.SDE_GENERIC
        if self.generic and inp.sde is None:
            raise ValueError("sde_kind GENERIC needs the user SDE object (Inputs.sde)")
        self.sde = inp.sde
        self.lib = _lib.load()
        dev = torch.device(device)
        self.dev = dev
        f = dict(device=dev, dtype=torch.float32)
        B, S = inp.x0.shape
        T, Cd = inp.context_full.shape[1] - 1, inp.context_full.shape[2]
        P, NL, H = inp.theta.shape[1], len(inp.w_hh), inp.w_hh[0].shape[1]
        self.B, self.T, self.S, self.C, self.P, self.H, self.NL = B, T, S, Cd, P, H, NL
        self.dims = _lib.Dims(B, T, S, Cd, P, H, NL, variant)
        self.dt, self.sde_kind, self.mask = float(inp.dt), inp.sde_kind, inp.positive_mask
        to = lambda t: t.to(dev).contiguous()  # noqa: E731
        self.x0, self.theta, self.eps = to(inp.x0), to(inp.theta), to(inp.eps)
        # bf16 = what the reference's autocast encoder hands to the head (grad_context comes back in the same dtype)
        self.ctx = inp.context_full.to(dev).to(context_dtype).contiguous()
        cdt = _lib.BF16 if context_dtype == torch.bfloat16 else _lib.F32
        # the 4*NL+2 parameter tensors are views of ONE flat buffer and so are their gradients (same offsets): the gradient
        # buffer is the all-reduce bucket and the pair is what the fused clip + AdamW + EMA step walks (make_optimizer)
        host_w = [list(ws) for ws in (inp.w_ih, inp.w_hh, inp.b_ih, inp.b_hh)]
        shapes = [tuple(t.shape) for ws in host_w for t in ws] + [tuple(inp.out_w.shape), tuple(inp.out_b.shape)]
        self.param_bucket = FlatBucket(shapes, dev)
        with torch.no_grad():
            for v, t in zip(self.param_bucket.views, [t for ws in host_w for t in ws] + [inp.out_w, inp.out_b]):
                v.copy_(t)
        pv = self.param_bucket.views
        self.w = [pv[i * NL:(i + 1) * NL] for i in range(4)]
        self.out_w, self.out_b = pv[4 * NL], pv[4 * NL + 1]
        self.bucket = FlatBucket(shapes, dev, extra=1)  # + the ELBO scalar: one all-reduce per iteration
        self.elbo_sign = torch.tensor([1.0, 1.0, -1.0, 1.0], **f) / B  # mean_b(obs + sde - gen + jac)
        v = self.bucket.views
        self.gw = [v[i * NL:(i + 1) * NL] for i in range(4)]
        self.g_out_w, self.g_out_b = v[4 * NL], v[4 * NL + 1]
        self.paths = torch.empty(B, T + 1, S, **f)
        self.means = torch.empty(B, T, S, **f)
        self.chol = torch.empty(B, T, S, S, **f)
        self.terms = torch.empty(B, 4, **f)
        s = -1.0 / B  # loss = -mean_b(obs + sde - gen + jac)
        self.g_terms = torch.tensor([s, s, -s, s], **f).repeat(B, 1).contiguous()
        self.g_z, self.g_means, self.g_chol = torch.empty_like(self.paths), torch.empty_like(self.means), torch.empty_like(self.chol)
        self.g_theta_elbo = torch.empty(B, P, **f)
        self.g_drift = torch.empty(B, T, S, **f) if self.generic else None
        self.g_diffusion = torch.empty(B, T, S, S, **f) if self.generic else None
        self.pos_dims = [s for s in range(S) if (inp.positive_mask >> s) & 1]
        self.grad_x0, self.grad_theta = torch.empty(B, S, **f), torch.empty(B, P, **f)
        self.grad_ctx = torch.zeros(B, T + 1, Cd, device=dev, dtype=context_dtype)
        u8 = dict(device=dev, dtype=torch.uint8)
        self.stash = torch.empty(self.lib.visde_stash_bytes(C.byref(self.dims)), **u8)
        self.ws_f = torch.empty(self.lib.visde_workspace_bytes(C.byref(self.dims), 0), **u8)
        self.ws_b = torch.empty(self.lib.visde_workspace_bytes(C.byref(self.dims), 1), **u8)
        self.obs_idx = inp.obs_idx.to(torch.int32).to(dev)
        self.obs_values = to(inp.obs_values)
        self.obs = _lib.Obs(self.obs_idx.shape[0], self.obs_values.shape[1], self.obs_idx.data_ptr(),
                            self.obs_values.data_ptr(), None, float(inp.obs_variance))
        self.cv = _lib.CtxView(self.ctx.data_ptr(), (T + 1) * Cd, Cd, cdt)
        self.gv = _lib.CtxView(self.grad_ctx.data_ptr(), (T + 1) * Cd, Cd, cdt)
        self.ws_struct = self._wstruct(self.w, self.out_w, self.out_b)
        self.gw_struct = self._wstruct(self.gw, self.g_out_w, self.g_out_b)
        # K0, K1, K5, K6, K2, grad_ctx, grad_theta, (2 NL + 1) x (split-K GEMM + reduce), theta-grad add
        self.launches_per_step = 7 + 2 * (2 * NL + 1) + 1

    def _wstruct(self, groups, ow, ob) -> _lib.Weights:
        s = _lib.Weights()
        for k in range(self.NL):
            s.w_ih[k], s.w_hh[k] = groups[0][k].data_ptr(), groups[1][k].data_ptr()
            s.b_ih[k], s.b_hh[k] = groups[2][k].data_ptr(), groups[3][k].data_ptr()
        s.out_w, s.out_b = ow.data_ptr(), ob.data_ptr()
        return s

    def head_weight_grads(self) -> List[Tensor]:
        return [*self.gw[0], *self.gw[1], *self.gw[2], *self.gw[3], self.g_out_w, self.g_out_b]

    def forward(self) -> None:
        lib, d = self.lib, self.dims
        st = torch.cuda.current_stream(self.dev).cuda_stream
        p = lambda t: t.data_ptr()  # noqa: E731
        _lib.check(lib.visde_path_fwd(C.byref(d), self.dt, p(self.x0), C.byref(self.cv), p(self.theta), p(self.eps),
                                      C.byref(self.ws_struct), p(self.paths), p(self.means), p(self.chol), p(self.stash),
                                      p(self.ws_f), self.ws_f.numel(), st))
        dr = df = None
        if self.generic:
            self._eval_user_sde()
            dr, df = p(self._drift_c), p(self._diffusion_c)
        _lib.check(lib.visde_elbo_fwd(C.byref(d), self.dt, self.sde_kind, self.mask, p(self.paths), p(self.means),
                                      p(self.chol), p(self.theta), dr, df, C.byref(self.obs), p(self.terms), st))

    def _eval_user_sde(self) -> None:
        """x_t = to_state(z_t) -> drift [B,T,S], diffusion [B,T,S,S] through the user's PyTorch callbacks, recorded for
        the vector-Jacobian product of the backward (evidence_lower_bound.py:31-40)."""
        B, T, S, P = self.B, self.T, self.S, self.P
        z_t = self.paths[:, :-1]
        x = z_t
        if self.pos_dims:
            x = z_t.clone()
            x[..., self.pos_dims] = torch.nn.functional.softplus(z_t[..., self.pos_dims])
        with torch.enable_grad():
            self._x_leaf = x.reshape(B * T, S).detach().requires_grad_(True)
            self._th_leaf = self.theta.detach().requires_grad_(True)
            th_flat = self._th_leaf[:, None, :].expand(B, T, P).reshape(B * T, P)
            self._drift = self.sde.drift(self._x_leaf, th_flat)
            self._diffusion = self.sde.diffusion(self._x_leaf, th_flat)
        self._drift_c = self._drift.detach().to(torch.float32).reshape(B, T, S).contiguous()
        self._diffusion_c = self._diffusion.detach().to(torch.float32).reshape(B, T, S, S).contiguous()

    def _user_sde_vjp(self) -> None:
        """g_drift / g_diffusion -> g_x (chained into g_z through the state transform) and g_theta."""
        B, T, S = self.B, self.T, self.S
        outs, gouts = [], []
        for o, g in ((self._drift, self.g_drift), (self._diffusion, self.g_diffusion)):
            if o.requires_grad:
                outs.append(o)
                gouts.append(g.reshape(o.shape).to(o.dtype))
        gx = gth = None
        if outs:
            gx, gth = torch.autograd.grad(outs, [self._x_leaf, self._th_leaf], gouts, allow_unused=True)
        if gx is not None:
            gx = gx.reshape(B, T, S)
            if self.pos_dims:
                gx = gx.clone()
                gx[..., self.pos_dims] *= torch.sigmoid(self.paths[:, :-1][..., self.pos_dims])
            self.g_z[:, :T].add_(gx)
        self._g_theta_sde = gth

    def backward(self) -> None:
        lib, d = self.lib, self.dims
        st = torch.cuda.current_stream(self.dev).cuda_stream
        p = lambda t: t.data_ptr()  # noqa: E731
        dr, df, gdr, gdf = ((p(self._drift_c), p(self._diffusion_c), p(self.g_drift), p(self.g_diffusion))
                            if self.generic else (None, None, None, None))
        _lib.check(lib.visde_elbo_bwd(C.byref(d), self.dt, self.sde_kind, self.mask, p(self.paths), p(self.means),
                                      p(self.chol), p(self.theta), dr, df, C.byref(self.obs), p(self.g_terms),
                                      p(self.g_z), p(self.g_means), p(self.g_chol), p(self.g_theta_elbo), gdr, gdf, st))
        if self.generic:
            self._user_sde_vjp()
        _lib.check(lib.visde_path_bwd(C.byref(d), self.dt, p(self.g_z), p(self.g_means), p(self.g_chol),
                                      C.byref(self.cv), p(self.theta), p(self.eps), C.byref(self.ws_struct),
                                      p(self.paths), p(self.stash), p(self.grad_x0), C.byref(self.gv), p(self.grad_theta),
                                      C.byref(self.gw_struct), p(self.ws_b), self.ws_b.numel(), st))
        self.grad_theta.add_(self.g_theta_elbo)
        if self.generic and self._g_theta_sde is not None:
            self.grad_theta.add_(self._g_theta_sde)

    def step(self) -> None:
        with torch.cuda.device(self.dev):
            self.forward()
            self.backward()

    def capture(self, post=None) -> None:
        """Capture one iteration (every kernel of K0..K4 + the ELBO kernels, all buffers static) into a CUDA graph:
        `replay()` then costs one launch instead of ~14 and leaves no gaps between the dependent kernels.  `post` (e.g. the
        NCCL all-reduce of the gradient bucket, which is capturable) is recorded behind the kernels in the same graph."""
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up on the capture side: per-kernel attributes, tensor-map encoders
            for _ in range(2):
                self.step()
                if post is not None:
                    post()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.step()
            if post is not None:
                post()

    def replay(self) -> None:
        self.graph.replay()

    def make_optimizer(self, lr: float = 1e-3, max_norm: float = 1.0, ema_decay: float | None = 0.999, **adamw):
        """Fused unscale + clip + AdamW + EMA over the head's flat parameter / gradient pair (inference/trainer.py:199-203,
        126): `opt.step()` after `replay()` (+ the all-reduce) is a complete data-parallel head iteration, no host sync."""
        from types import SimpleNamespace

        from viforsdes_b200.optim import FusedAdamWEma

        n = self.param_bucket.flat.numel()
        flat = SimpleNamespace(data=self.param_bucket.flat, grads=self.bucket.flat[:n], segments=[(0, n)],
                               groups=[list(self.param_bucket.views)])
        return FusedAdamWEma(flat, [lr], max_norm=max_norm, ema_decay=ema_decay, **adamw)

    def stage_elbo(self) -> None:
        """Write the batch-mean ELBO of this rank into the bucket's tail slot (two tiny kernels)."""
        torch.sum(torch.mv(self.terms, self.elbo_sign), dim=0, keepdim=True, out=self.bucket.extra)

    def results(self) -> Dict[str, object]:
        grads = {"x0": self.grad_x0, "context": self.grad_ctx[:, : self.T], "theta": self.grad_theta,
                 "out_w": self.g_out_w, "out_b": self.g_out_b}
        for k in range(self.NL):
            grads[f"w_ih_l{k}"], grads[f"w_hh_l{k}"] = self.gw[0][k], self.gw[1][k]
            grads[f"b_ih_l{k}"], grads[f"b_hh_l{k}"] = self.gw[2][k], self.gw[3][k]
        return {"paths": self.paths, "means": self.means, "chol": self.chol, "terms": self.terms, "grads": grads}
