"""Golden vectors for the callers either side of the path, from the REAL reference.

    python tests/golden/make_golden_aux.py        (build container only: imports /root/reference/src)

  em_{ou,lv}.pt : core/euler_maruyama.py ``euler_maruyama`` with injected noise on the example SDEs
                  (examples/ornstein_uhlenbeck.py, examples/lotka_volterra.py), the pre-training objective of
                  inference/trainer.py:253-259 on it and its autograd gradient w.r.t. theta;
  summary_lv.pt : ``StateSpace.to_state`` + mean / std over samples (posterior/variational_posterior.py:116-135);
  ema.pt        : ``ExponentialMovingAverage.update`` driven by torch.optim.AdamW + clip_grad_norm_
                  (inference/trainer.py:199-203,126) on a small module.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(REF / "src"))
for name, attrs in {"matplotlib": {}, "matplotlib.pyplot": {}, "matplotlib.axes": {"Axes": object},
                    "matplotlib.figure": {"Figure": object}}.items():
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)

import torch  # noqa: E402
from torch import nn  # noqa: E402

from variational_sde.core.euler_maruyama import euler_maruyama  # noqa: E402
from variational_sde.inference.exponential_moving_average import ExponentialMovingAverage  # noqa: E402
from variational_sde.inference.state_space import StateSpace  # noqa: E402


def _load_example(name: str):
    spec = importlib.util.spec_from_file_location(name, REF / "examples" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def em_case(kind: str, B: int, horizon: float, dt: float, seed: int) -> dict:
    g = torch.Generator().manual_seed(seed)
    if kind == "ou":
        sde = _load_example("ornstein_uhlenbeck").OrnsteinUhlenbeck()
        theta = torch.stack([torch.exp(0.3 * torch.randn(B, generator=g)), 1 + 0.3 * torch.randn(B, generator=g),
                             torch.exp(-1 + 0.3 * torch.randn(B, generator=g))], 1)
        pos, obs_values = [], torch.tensor([[0.2], [0.9], [1.1], [0.7]])
    else:
        sde = _load_example("lotka_volterra").LotkaVolterra()
        theta = torch.exp(torch.log(torch.tensor([0.5, 0.0025, 0.3])) + 0.1 * torch.randn(B, 3, generator=g))
        pos, obs_values = [0, 1], torch.tensor([[71.0, 79.0], [120.0, 60.0], [160.0, 140.0], [60.0, 200.0]])
    # a few LV trajectories start next to the clamp so that the clamp(min=1e-6) branch is exercised
    S = obs_values.shape[1]
    n_steps = round(horizon / dt)
    obs_times = torch.linspace(0, horizon, obs_values.shape[0])
    noise = torch.randn(B, n_steps, S, generator=g)
    x0 = obs_values[0].unsqueeze(0).expand(B, -1).clone()
    if kind == "lv":
        x0[-2:] = torch.tensor([[2e-6, 3.0], [1.5, 1e-6]])
        noise[-2:, :3] = -3.0
    theta = theta.requires_grad_(True)
    x0 = x0.requires_grad_(True)
    paths = euler_maruyama(sde, x0, theta, horizon, dt, pos, noise=noise)
    obs_idx = (obs_times / dt).round().long()
    mse = ((paths[:, obs_idx] - obs_values) ** 2).mean()
    mse.backward()
    return {"kind": kind, "dt": dt, "horizon": horizon, "positive_dims": pos, "x0": x0.detach(), "theta": theta.detach(),
            "noise": noise, "obs_times": obs_times, "obs_values": obs_values, "paths": paths.detach(), "mse": mse.detach(),
            "g_theta": theta.grad.clone(), "g_x0": x0.grad.clone()}


def summary_case() -> dict:
    g = torch.Generator().manual_seed(5)
    z = 3.0 * torch.randn(37, 21, 2, generator=g)
    z[0, 0, 0] = 25.0  # above F.softplus's threshold
    ss = StateSpace(2, [1])
    x = ss.to_state(z)
    return {"z": z, "positive_dims": [1], "x": x, "mean": x.mean(dim=0), "std": x.std(dim=0)}


def ema_case() -> dict:
    torch.manual_seed(3)
    model = nn.Sequential(nn.Linear(5, 7), nn.Tanh(), nn.Linear(7, 3))
    init = [p.detach().clone() for p in model.parameters()]
    ema = ExponentialMovingAverage(model, decay=0.9)
    groups = [{"params": list(model[0].parameters()), "lr": 1e-2}, {"params": list(model[2].parameters()), "lr": 3e-3}]
    opt = torch.optim.AdamW(groups)
    g = torch.Generator().manual_seed(4)
    grads, norms = [], []
    for _ in range(6):
        gs = [2.0 * torch.randn(p.shape, generator=g) for p in model.parameters()]
        grads.append(gs)
        for p, gr in zip(model.parameters(), gs):
            p.grad = gr.clone()
        norms.append(nn.utils.clip_grad_norm_(model.parameters(), 1.0))
        opt.step()
        ema.update()
    return {"init": init, "grads": grads, "lrs": [1e-2, 1e-2, 3e-3, 3e-3], "max_norm": 1.0, "decay": 0.9,
            "params": [p.detach().clone() for p in model.parameters()], "shadow": list(ema.state_dict().values()),
            "norms": torch.stack(norms)}


def main() -> None:
    torch.save(em_case("ou", 6, 2.0, 0.05, 11), HERE / "em_ou.pt")
    torch.save(em_case("lv", 6, 2.0, 0.05, 12), HERE / "em_lv.pt")
    torch.save(summary_case(), HERE / "summary_lv.pt")
    torch.save(ema_case(), HERE / "ema.pt")
    print("wrote em_ou em_lv summary_lv ema")


if __name__ == "__main__":
    main()
