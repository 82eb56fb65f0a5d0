"""Small launches of every kernel family for `compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_smoke.py`
(the reference has no race detection of its own; SURVEY.md lists it among the auxiliary subsystems to cover)."""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import oracle_torch as O  # noqa: E402  (problem generator only; nothing is checked against it here)
from tests._util import build_head, cuda_inputs  # noqa: E402
from viforsdes_b200 import _lib, ops, sde as vs  # noqa: E402
from viforsdes_b200.euler_maruyama import euler_maruyama  # noqa: E402
from viforsdes_b200.optim import FlatParameters, FusedAdamWEma  # noqa: E402
from viforsdes_b200.posterior import summarise_paths  # noqa: E402
from viforsdes_b200.state_space import StateSpace  # noqa: E402


def path_case(kind, B, T, variant, **kw):
    ops.set_variant(variant)
    p = O.make_problem(kind, B, T, **kw)
    head = build_head(p)
    x0, full, view, theta, eps = cuda_inputs(p)
    out = head.sample_diffusion_paths(x0, view, theta, eps, p.dt)
    sum(o.sum() for o in out).backward()
    torch.cuda.synchronize()
    print("path", kind, B, T, variant, "ok")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "path"):
        path_case("lv", 5, 9, _lib.VARIANT_FAST, context_dim=16, hidden_dim=64, num_layers=2)
        path_case("lv", 37, 7, _lib.VARIANT_TILED, context_dim=16, hidden_dim=64, num_layers=2)   # 4-trajectory tiles, ragged
        path_case("l96", 601, 3, _lib.VARIANT_TILED, context_dim=8, hidden_dim=64, num_layers=2, state_dim=4)  # 8-trajectory tiles
        path_case("l96", 3, 5, _lib.VARIANT_FAST, context_dim=8, hidden_dim=64, num_layers=2, state_dim=10)     # wide-state family
        path_case("ou", 3, 5, _lib.VARIANT_GENERIC, context_dim=8, hidden_dim=48, num_layers=3)
        path_case("lv", 130, 4, _lib.VARIANT_TC, context_dim=128, hidden_dim=64, num_layers=2)                  # tcgen05 recurrence
        ops.set_variant(_lib.VARIANT_AUTO)
    if which in ("all", "aux"):
        B, T = 70, 45
        th = torch.tensor([[0.5, 0.0025, 0.3]]).repeat(B, 1).cuda().requires_grad_(True)
        x0 = torch.tensor([[71.0, 79.0]]).repeat(B, 1).cuda()
        for noise in (None, torch.randn(B, T, 2, device="cuda")):
            p = euler_maruyama(vs.LotkaVolterra(), x0, th, T * 0.05, 0.05, [0, 1], noise=noise, seed=3)
            p.sum().backward()
        torch.cuda.synchronize()
        print("em ok")
        summarise_paths(torch.randn(100, 33, 3, device="cuda"), StateSpace(3, [1]))
        torch.cuda.synchronize()
        print("summary ok")
        par = torch.nn.Parameter(torch.randn(4099, device="cuda"))
        flat = FlatParameters([[par]])
        opt = FusedAdamWEma(flat, lrs=[1e-3], max_norm=1.0, ema_decay=0.99)
        par.grad.copy_(torch.randn(4099, device="cuda"))
        opt.step()
        torch.cuda.synchronize()
        print("adamw ok")


if __name__ == "__main__":
    main()
