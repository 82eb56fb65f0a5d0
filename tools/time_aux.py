"""Device timings of the callers either side of the path on one B200, next to what the reference does for
the same step (its PyTorch loop / foreach optimiser on the same GPU) -- CUDA events, warm-up, median of `reps`.

    python tools/time_aux.py            -> one JSON line per component (gpurun_out/time_aux.jsonl when the directory exists)
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch
from torch import nn

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from viforsdes_b200 import sde as vs  # noqa: E402
from viforsdes_b200.euler_maruyama import _mask, euler_maruyama, philox_normal  # noqa: E402
from viforsdes_b200.optim import FlatParameters, FusedAdamWEma  # noqa: E402
from viforsdes_b200.posterior import summarise_paths  # noqa: E402
from viforsdes_b200.state_space import StateSpace  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def reference_em(sde, x0, theta, n_steps, dt, pos, noise):
    """core/euler_maruyama.py:27-45 as the reference runs it (PyTorch ops, one round per step)."""
    sq = dt**0.5
    traj = torch.empty(x0.shape[0], n_steps + 1, x0.shape[1], device=x0.device)
    traj[:, 0] = x0
    x = x0.clone()
    for step in range(n_steps):
        x = x + sde.drift(x, theta) * dt + torch.einsum("bij,bj->bi", sde.diffusion(x, theta), noise[:, step]) * sq
        if pos:
            x[:, pos] = x[:, pos].clamp(min=1e-6)
        traj[:, step + 1] = x
    return traj


def main():
    out = []
    dev = "cuda"
    # --- pre-training simulator: inference/trainer.py:208-259 at the Lotka-Volterra example's sizes
    B, T, dt = 4096, 800, 0.05
    g = torch.Generator().manual_seed(0)
    theta0 = torch.exp(torch.log(torch.tensor([0.5, 0.0025, 0.3])) + 0.1 * torch.randn(B, 3, generator=g)).to(dev)
    x0 = torch.tensor([71.0, 79.0]).expand(B, 2).contiguous().to(dev)
    obs_idx = torch.tensor([0, 200, 400, 600, 800], device=dev)
    obs_v = torch.tensor([[71.0, 79.0], [120.0, 60.0], [160.0, 140.0], [60.0, 200.0], [50.0, 90.0]], device=dev)
    sde, pos = vs.LotkaVolterra(), [0, 1]
    noise = philox_normal(1, B, T, 2)

    def fused(noise_arg):
        th = theta0.clone().requires_grad_(True)
        p = euler_maruyama(sde, x0, th, T * dt, dt, pos, noise=noise_arg, seed=1)
        ((p[:, obs_idx] - obs_v) ** 2).mean().backward()
        return th.grad

    def ref():
        th = theta0.clone().requires_grad_(True)
        p = reference_em(sde, x0, th, T, dt, pos, noise)
        ((p[:, obs_idx] - obs_v) ** 2).mean().backward()
        return th.grad

    g_f, g_r = fused(noise), ref()
    t_philox, t_inj, t_ref = timed(lambda: fused(None)), timed(lambda: fused(noise)), timed(ref, reps=3, warm=1)
    # kernels alone (no autograd glue / gather): forward + reverse
    kind, mask = sde.device_kind, _mask(pos)
    th = theta0.clone().requires_grad_(True)
    gp = torch.zeros(B, T + 1, 2, device=dev)
    gp[:, obs_idx] = 1e-3

    def kernels():
        p = torch.ops.visde.em_fwd(x0, th, None, 1, T, dt, kind, mask)
        torch.ops.visde.em_bwd(gp, p, th, None, 1, dt, kind, mask)

    t_k = timed(kernels)
    bytes_k = 4 * B * (T + 1) * 2 * 3  # paths written, paths + cotangents read
    out.append({"component": "euler_maruyama fwd+bwd (pretrain objective)", "workload": f"lv B={B} T={T}",
                "fused_philox_ms": t_philox, "fused_injected_noise_ms": t_inj, "kernels_only_ms": t_k,
                "kernels_only_gbs": bytes_k / t_k / 1e6, "reference_pytorch_loop_same_gpu_ms": t_ref,
                "speedup_vs_reference_loop": t_ref / t_philox,
                "grad_theta_normwise_diff_vs_loop": ((g_f - g_r).abs().max() / g_r.abs().max()).item(),
                "traj_steps_per_s": B * T / (t_philox * 1e-3)})

    # --- posterior summary: variational_posterior.py:116-135 (n_samples = 1000) and a large-n case
    for n, T1, S, posd in ((1000, 801, 2, [0, 1]), (65536, 101, 2, [0, 1])):
        z = 3 * torch.randn(n, T1, S, device=dev)
        ss = StateSpace(S, posd)

        def ref_sum():
            x = ss.to_state(z)
            return x, x.mean(dim=0), x.std(dim=0)

        t_f, t_r = timed(lambda: summarise_paths(z, ss)), timed(ref_sum)
        nbytes = 4 * n * T1 * S * 2
        out.append({"component": "posterior summary (from_latent + mean + std)", "workload": f"n={n} T+1={T1} S={S}",
                    "fused_ms": t_f, "fused_gbs": nbytes / t_f / 1e6, "reference_pytorch_same_gpu_ms": t_r,
                    "speedup": t_r / t_f})

    # --- optimiser tail: trainer.py:199-203,126 on the example model's size (8.28 M fp32 parameters) and the head alone
    for n in (8_280_000, 87_750):
        p_f = nn.Parameter(torch.randn(n, device=dev))
        flat = FlatParameters([[p_f]])
        opt = FusedAdamWEma(flat, lrs=[1e-3], max_norm=1.0, ema_decay=0.999)
        p_f.grad.copy_(torch.randn(n, device=dev))
        # reference: ~100 parameter tensors; split the same memory into 100 chunks for the foreach path
        chunks = [nn.Parameter(c.clone()) for c in torch.randn(n, device=dev).chunk(100)]
        ref_opt = torch.optim.AdamW(chunks, lr=1e-3)
        shadow = [c.detach().clone() for c in chunks]
        for c in chunks:
            c.grad = torch.randn_like(c)

        def ref_step():
            nn.utils.clip_grad_norm_(chunks, 1.0)
            ref_opt.step()
            with torch.no_grad():
                for s, c in zip(shadow, chunks):
                    s.lerp_(c.detach(), 1e-3)

        t_f, t_r = timed(opt.step), timed(ref_step)
        nbytes = 4 * n * (1 + 5 + 4)  # norm pass reads g; update reads p, g, m, v, ema and writes p, m, v, ema
        out.append({"component": "clip + AdamW + EMA", "workload": f"n={n} fp32", "fused_ms": t_f, "fused_gbs": nbytes / t_f / 1e6,
                    "reference_foreach_100_tensors_same_gpu_ms": t_r, "speedup": t_r / t_f})

    lines = [json.dumps(o) for o in out]
    print("\n".join(lines))
    d = ROOT / "gpurun_out"
    if d.exists():
        (d / "time_aux.jsonl").write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
