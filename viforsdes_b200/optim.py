"""Fused optimiser tail of a training iteration (inference/trainer.py:199-203 + :126): unscale, global-norm clip,
AdamW and the EMA update over ONE flat fp32 buffer in two kernel launches and no host sync (csrc/optim.cu).

``FlatParameters`` re-points the parameters (and gradients) of a module at slices of flat buffers, so the gradient
buffer is at the same time the NCCL all-reduce bucket (``dist.allreduce_mean_(flat.grads)``).  Parameter groups with
their own learning rate (training_context.py:97-102: ``learning_rate`` / ``sde_param_lr``) are contiguous segments."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, Iterable, Iterator, List, Optional, Sequence

import torch
from torch import Tensor, nn

from viforsdes_b200 import _lib
from viforsdes_b200.ops import _ptr, _require_cuda, _stream


class FlatParameters:
    def __init__(self, groups: Sequence[Iterable[nn.Parameter]]) -> None:
        self.groups: List[List[nn.Parameter]] = [list(g) for g in groups]
        params = [p for g in self.groups for p in g]
        if not params:
            raise ValueError("no parameters")
        dev = params[0].device
        if any(p.dtype != torch.float32 or p.device != dev for p in params):
            raise ValueError("FlatParameters needs fp32 parameters on one device (master weights)")
        total, self.segments = 0, []
        offs = []
        for g in self.groups:
            start = total
            for p in g:
                offs.append(total)
                total += (p.numel() + 3) // 4 * 4  # every tensor 16-byte aligned
            self.segments.append((start, total))
        self.data = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grads = torch.zeros(total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(params, offs):
                n = p.numel()
                self.data[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.data[o:o + n].view_as(p)
                p.grad = self.grads[o:o + n].view_as(p)
        self.params = params

    def zero_grad(self) -> None:
        """Keeps the gradient views attached (``set_to_none`` would detach them from the bucket)."""
        self.grads.zero_()


class FusedAdamWEma:
    """torch.optim.AdamW defaults (betas (0.9, 0.999), eps 1e-8, weight_decay 1e-2) + clip_grad_norm_ + EMA lerp."""

    def __init__(self, flat: FlatParameters, lrs: Sequence[float], betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, max_norm: float = 0.0, ema_decay: Optional[float] = None) -> None:
        if len(lrs) != len(flat.segments):
            raise ValueError("one learning rate per parameter group")
        _require_cuda(flat.data)
        self.flat, self.lrs, self.betas, self.eps, self.wd = flat, list(lrs), betas, eps, weight_decay
        self.max_norm, self.ema_decay = max_norm, ema_decay
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat.data), torch.zeros_like(flat.data)
        self.ema = flat.data.clone() if ema_decay is not None else None
        self.sqnorm = torch.zeros(1, device=flat.data.device, dtype=torch.float32)
        # steps skipped because a gradient was inf / NaN (GradScaler.step semantics, trainer.py:202), counted on the device
        self.skipped_steps = torch.zeros(1, device=flat.data.device, dtype=torch.int64)
        self.lib = _lib.load()
        self.ws = torch.empty(self.lib.visde_grad_sqnorm_workspace_bytes(), device=flat.data.device, dtype=torch.uint8)
        self.step_count = 0

    @property
    def grad_norm(self) -> Tensor:
        """Device scalar (pre-clip global norm, what ``clip_grad_norm_`` returns); reading it is the caller's sync."""
        return self.sqnorm.sqrt().squeeze(0)

    @property
    def found_inf(self) -> Tensor:
        """Device bool: the last step saw a non-finite gradient and was skipped (what ``scaler.update()`` needs)."""
        return ~torch.isfinite(self.sqnorm).squeeze(0)

    def _check_views(self) -> None:
        """The parameters / gradients must still alias the flat buffers: ``zero_grad(set_to_none=True)`` or a
        ``module.to()`` after flattening silently detaches them (the fused step would then read stale zeros)."""
        f = self.flat
        params = getattr(f, "params", None)
        if not params:
            return
        base_p, base_g, end_p = f.data.data_ptr(), f.grads.data_ptr(), f.data.data_ptr() + 4 * f.data.numel()
        for p in params:
            if not (base_p <= p.data_ptr() < end_p):
                raise RuntimeError("a parameter no longer aliases FlatParameters.data (module.to()/.float() after flattening?)")
            off = p.data_ptr() - base_p
            if p.grad is None or p.grad.data_ptr() != base_g + off:
                p.grad = f.grads[off // 4: off // 4 + p.numel()].view_as(p)  # re-attach (zero_grad(set_to_none=True))

    def step(self, inv_scale: Optional[Tensor] = None) -> None:
        """One fused unscale + clip + AdamW + EMA step.  With `inv_scale` (a GradScaler is in use) or clipping the
        squared gradient norm is computed first; if it is non-finite the update is skipped on the device exactly like
        ``scaler.step(optimizer)`` (parameters, moments and the bias-correction step count stay; the EMA still moves)."""
        f, lib = self.flat, self.lib
        self._check_views()
        self.step_count += 1
        st = _stream()
        with torch.cuda.device(f.data.device):
            clip = self.max_norm > 0
            guard = clip or inv_scale is not None
            if guard:
                _lib.check(lib.visde_grad_sqnorm(f.grads.numel(), _ptr(f.grads), _ptr(inv_scale), 0, _ptr(self.sqnorm),
                                                 _ptr(self.skipped_steps), _ptr(self.ws), self.ws.numel(), st))
            for (lo, hi), lr in zip(f.segments, self.lrs):
                if hi == lo:
                    continue
                sl = slice(lo, hi)
                _lib.check(lib.visde_adamw_ema_step(
                    hi - lo, _ptr(f.data[sl]), _ptr(f.grads[sl]), _ptr(self.exp_avg[sl]), _ptr(self.exp_avg_sq[sl]),
                    _ptr(self.ema[sl]) if self.ema is not None else None, lr, self.betas[0], self.betas[1], self.eps,
                    self.wd, self.step_count, self.max_norm if clip else 0.0, _ptr(self.sqnorm) if guard else None,
                    _ptr(inv_scale), self.ema_decay if self.ema_decay is not None else 0.0,
                    _ptr(self.skipped_steps) if guard else None, st))

    @contextmanager
    def ema_applied(self) -> Iterator[None]:
        """``ExponentialMovingAverage.apply`` (exponential_moving_average.py:30-42): run the body with the shadow weights in
        place of the parameters (posterior sampling), then restore them.  Two flat copies instead of 2 x #tensors."""
        if self.ema is None:
            raise RuntimeError("this optimiser keeps no EMA shadow (ema_decay=None)")
        backup = self.flat.data.clone()
        with torch.no_grad():
            self.flat.data.copy_(self.ema)
        try:
            yield
        finally:
            with torch.no_grad():
                self.flat.data.copy_(backup)

    def state_dict(self) -> Dict[str, object]:
        """Checkpoint / resume: moments, shadow and step count (parameters live in the module's own state_dict)."""
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "ema": None if self.ema is None else self.ema.clone(),
                "step": self.step_count - int(self.skipped_steps.item())}

    def load_state_dict(self, state: Dict[str, object]) -> None:
        n = self.exp_avg.numel()
        if state["exp_avg"].numel() != n or state["exp_avg_sq"].numel() != n:
            raise ValueError("optimizer state does not match the flat parameter buffer")
        if self.ema is not None:
            if state.get("ema") is None or state["ema"].numel() != n:
                raise ValueError("checkpoint has no EMA shadow of the right size for this optimiser (ema_decay is set)")
            self.ema.copy_(state["ema"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        self.step_count = int(state["step"])
        self.skipped_steps.zero_()

    def ema_views(self) -> List[Tensor]:
        """EMA shadow tensors shaped like the parameters (``ExponentialMovingAverage.shadow`` values, in order)."""
        if self.ema is None:
            return []
        out, off = [], 0
        for g in self.flat.groups:
            for p in g:
                out.append(self.ema[off:off + p.numel()].view_as(p))
                off += (p.numel() + 3) // 4 * 4
        return out
