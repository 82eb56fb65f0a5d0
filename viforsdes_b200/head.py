"""``DiffusionTransitionHead`` -- drop-in for src/variational_sde/models/head.py:20-209.

Same constructor, parameter / buffer names (so reference checkpoints load: ``gru.weight_ih_l{k}``,
``gru.weight_hh_l{k}``, ``gru.bias_ih_l{k}``, ``gru.bias_hh_l{k}``, ``out_proj.weight``,
``out_proj.bias``, ``_tril_rows``, ``_tril_cols``, ``_diag_mask``) and the same
``HeadProtocol.sample_diffusion_paths`` signature (inference/diffusion_path_sampler.py:24-32).
Only the dispatch changes: the Triton ``_SDEFunction`` / ``launch_fwd`` are replaced by the
``visde::path_fwd`` custom op (sm_100a CUDA kernels); the nn.GRU weights are read in their native
layout, so the per-call stack / transpose copies of head.py:106-154 and weights.py:176-196 vanish.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor, nn

from viforsdes_b200 import ops  # noqa: F401  (registers torch.ops.visde.*)

MAX_LAYERS = 4  # src/variational_sde/kernels/constants.py:13
DIAG_MIN = 1e-2  # src/variational_sde/inference/constants.py:6


@dataclass(frozen=True)
class HeadConfig:
    """src/variational_sde/config.py:83-92."""

    hidden_dim: int = 64
    num_layers: int = 2

    def __post_init__(self) -> None:
        if self.hidden_dim <= 0 or self.num_layers <= 0:
            raise ValueError("value must be positive")


class _FlooredDiagonal(torch.autograd.Function):
    """``lower_bound`` of src/variational_sde/primitives/bounds.py:10-31: max(x, floor) whose cotangent passes where the
    input is above the floor OR the cotangent is negative (pushing a clamped entry back up stays possible).  The kernels
    bake the same rule in (csrc/path_*.cu); this PyTorch form serves the single-step ``forward``."""

    @staticmethod
    def forward(ctx, x: Tensor, floor: float) -> Tensor:  # type: ignore[override]
        ctx.save_for_backward(x)
        ctx.floor = floor
        return x.clamp_min(floor)

    @staticmethod
    def backward(ctx, g: Tensor):  # type: ignore[override]
        (x,) = ctx.saved_tensors
        return torch.where((x >= ctx.floor) | (g < 0), g, torch.zeros_like(g)), None


class DiffusionTransitionHead(nn.Module):
    _tril_rows: Tensor
    _tril_cols: Tensor
    _diag_mask: Tensor

    def __init__(self, state_dim: int, context_dim: int, sde_param_dim: int, config: HeadConfig) -> None:
        super().__init__()
        if config.num_layers < 1 or config.num_layers > MAX_LAYERS:
            raise ValueError(f"num_layers must be in [1, {MAX_LAYERS}], got {config.num_layers}")
        self.state_dim = state_dim
        self.context_dim = context_dim
        self.sde_param_dim = sde_param_dim
        self.hidden_dim = config.hidden_dim
        self.num_layers = config.num_layers
        self.n_tril = state_dim * (state_dim + 1) // 2

        rows, cols = torch.tril_indices(state_dim, state_dim)
        self.register_buffer("_tril_rows", rows)
        self.register_buffer("_tril_cols", cols)
        self.register_buffer("_diag_mask", rows == cols)

        self.gru = nn.GRU(input_size=state_dim + context_dim + sde_param_dim, hidden_size=config.hidden_dim,
                          num_layers=config.num_layers, batch_first=True)
        self.out_proj = nn.Linear(config.hidden_dim, state_dim + self.n_tril)
        self._init_out_proj()

    def _init_out_proj(self) -> None:
        # head.py:60-66: mu = 0 and L = I at initialisation
        with torch.no_grad():
            self.out_proj.weight.zero_()
            self.out_proj.bias.zero_()
            for k in range(self.state_dim):
                self.out_proj.bias[self.state_dim + k * (k + 3) // 2] = 1.0

    def init_hidden(self, batch: int, device: torch.device, dtype: torch.dtype = torch.float32) -> Tensor:
        return torch.zeros(self.num_layers, batch, self.hidden_dim, device=device, dtype=dtype)

    def forward(self, x_t: Tensor, context_t: Tensor, sde_parameters: Tensor,
                hidden: Tensor | None = None) -> tuple[Tensor, Tensor, Tensor]:
        """The single-step spec of the transition (src/variational_sde/models/head.py:68-86), in plain PyTorch on any
        device: one ``nn.GRU`` step on cat[x_t, context_t, theta] -> out_proj -> (mu [B,S], L [B,S,S], hidden [NL,B,H]).
        Not the hot path -- ``sample_diffusion_paths`` runs all T steps inside the fused kernels -- but the API users step
        manually, and the statement of the step math the kernels are tested against."""
        step_in = torch.cat([x_t, context_t, sde_parameters], dim=-1).unsqueeze(1)
        out, hidden = self.gru(step_in, hidden)
        params = self.out_proj(out.squeeze(1))
        return params[..., : self.state_dim], self._tril_from_params(params[..., self.state_dim:]), hidden

    def _tril_from_params(self, params: Tensor) -> Tensor:
        """Row-major lower-triangular fill, diagonal floored at DIAG_MIN (head.py:88-97)."""
        vals = torch.where(self._diag_mask, _FlooredDiagonal.apply(params, DIAG_MIN), params)
        L = params.new_zeros(params.shape[0], self.state_dim, self.state_dim)
        L[:, self._tril_rows, self._tril_cols] = vals
        return L

    def _weight_lists(self):
        nl = self.num_layers
        g = self.gru
        return ([getattr(g, f"weight_ih_l{k}") for k in range(nl)], [getattr(g, f"weight_hh_l{k}") for k in range(nl)],
                [getattr(g, f"bias_ih_l{k}") for k in range(nl)], [getattr(g, f"bias_hh_l{k}") for k in range(nl)])

    def sample_diffusion_paths(self, x0: Tensor, context: Tensor, sde_parameters: Tensor, standard_noise: Tensor,
                               time_step: float) -> tuple[Tensor, Tensor, Tensor]:
        """x0 [B,S] latent, context [B,T,C] (fp32/bf16, may be the strided view context[:, :-1]),
        sde_parameters [B,P], standard_noise [B,T,S] -> paths [B,T+1,S], means [B,T,S], chol [B,T,S,S].
        Training mode records the autograd graph (head.py:164-199); eval mode is the no-grad,
        no-stash launch (head.py:200-209)."""
        if context.shape[-1] != self.context_dim or sde_parameters.shape[-1] != self.sde_param_dim:
            raise ValueError("context / sde_parameters feature dims do not match the head")
        w_ih, w_hh, b_ih, b_hh = self._weight_lists()
        if self.training and torch.is_grad_enabled():
            paths, means, chol, _ = torch.ops.visde.path_fwd(
                x0, context, sde_parameters, standard_noise, w_ih, w_hh, b_ih, b_hh, self.out_proj.weight,
                self.out_proj.bias, float(time_step), True)
        else:
            with torch.no_grad():
                paths, means, chol, _ = torch.ops.visde.path_fwd(
                    x0, context, sde_parameters, standard_noise, w_ih, w_hh, b_ih, b_hh, self.out_proj.weight,
                    self.out_proj.bias, float(time_step), False)
        if x0.dtype != torch.float32:  # autograd.py:117-122
            return paths.to(x0.dtype), means.to(x0.dtype), chol.to(x0.dtype)
        return paths, means, chol

    def sample_diffusion_paths_from_tokens(self, x0: Tensor, tokens: Tensor, output_proj: nn.Linear,
                                           sde_parameters: Tensor, standard_noise: Tensor,
                                           time_step: float) -> tuple[Tensor, Tensor, Tensor]:
        """Context-producer fusion.  The encoder ends in a plain ``output_proj`` Linear
        (primitives/sit.py:156-158,185) and the head starts with the context columns of ``weight_ih_l0``: two
        back-to-back linear maps.  Fold them, ``W' = W_ih[:, ctx] . W_op`` and ``b' = b_ih + W_ih[:, ctx] . b_op``
        (a [3H, C] x [C, E] product per step), and feed the encoder's pre-projection ``tokens`` [B, T, E] to the
        same kernels with C := E: the [B, T, C] context tensor is never written or read, the encoder's output GEMM
        and its backward disappear, and autograd distributes the folded gradients to both weight matrices."""
        S, Cd = self.state_dim, self.context_dim
        if output_proj.out_features != Cd or tokens.shape[-1] != output_proj.in_features:
            raise ValueError("output_proj must map the token width to the head's context_dim")
        w_ih, w_hh, b_ih, b_hh = self._weight_lists()
        w0 = w_ih[0]
        w_ctx = w0[:, S:S + Cd]
        folded = torch.cat([w0[:, :S], w_ctx @ output_proj.weight.to(w0.dtype), w0[:, S + Cd:]], dim=1)
        b0 = b_ih[0] if output_proj.bias is None else b_ih[0] + w_ctx @ output_proj.bias.to(w0.dtype)
        save = self.training and torch.is_grad_enabled()
        args = (x0, tokens, sde_parameters, standard_noise, [folded, *w_ih[1:]], w_hh, [b0, *b_ih[1:]], b_hh,
                self.out_proj.weight, self.out_proj.bias, float(time_step), save)
        if save:
            paths, means, chol, _ = torch.ops.visde.path_fwd(*args)
        else:
            with torch.no_grad():
                paths, means, chol, _ = torch.ops.visde.path_fwd(*args)
        if x0.dtype != torch.float32:
            return paths.to(x0.dtype), means.to(x0.dtype), chol.to(x0.dtype)
        return paths, means, chol
