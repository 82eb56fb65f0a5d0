#!/usr/bin/env python
"""Benchmark of the variational path-sampling hot path (BASELINE.json metric: trajectory-steps/s
for fwd + bwd ELBO, ms per ELBO iteration).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A step = one pass of the hot path over one batch of synthetic input:
  K0 ctx GEMM -> K1 path fwd -> K5 ELBO fwd -> K6 ELBO bwd -> K2 path bwd -> K3 grad_ctx -> K4 wgrads
  (+ for N > 1: ONE NCCL average of the head's flat weight-gradient bucket with the ELBO scalar in its tail slot).
`value`   : inputs resident in HBM, CUDA-event time per step on the launching stream, max over ranks.
`e2e`     : the same iteration through the host-buffer C-ABI entry (visde_session_step): pinned HOST
            buffers in, H2D + kernels + D2H of ELBO terms and gradients inside the timed region.
`roofline`: dominant kernel (by device time in the timed region, from the library's stage profiler).
`cpu_baseline`: the CPU oracle (stepwise PyTorch port of the reference path) on the host cores.
`--impl reference` times that CPU port alone (rank 0 only) and prints the same JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "trajectory-steps/s (fwd+bwd ELBO)"
UNIT = "trajectory-steps/s"
DEFAULT_WORKLOAD = "lv_b128_t800"  # BASELINE.json configs[1]: Lotka-Volterra, dt=0.05, batch 128, 1 B200
REFERENCE_SAMPLE_STEPS = 200       # --impl reference: first 200 of the T grid steps per timed step


def path_bytes_per_unit(S: int, Cd: int) -> int:
    """operator-contract HBM bytes per trajectory-step, fp32 I/O, stash excluded."""
    return 4 * (3 * Cd + 13 * S + 5 * S * S)


def path_flops_per_unit(S: int, Cd: int, H: int, NL: int) -> int:
    G, nt = 3 * H, S * (S + 1) // 2
    mac = (S + Cd) * G + H * G + (NL - 1) * 2 * H * G + H * (S + nt) + nt
    return 6 * mac


# per-kernel ALGORITHMIC bytes per trajectory-step (what the op contract + the hoisting design make
# each kernel move at minimum; stated in DESIGN.md §5)
def stage_bytes_per_unit(stage: str, S: int, Cd: int, H: int, NL: int) -> int:
    nt, G = S * (S + 1) // 2, 3 * H
    stash = NL * 5 * H + nt
    dg = NL * 4 * H + S + nt
    return 4 * {
        "K0_ctx_gemm": Cd + G,                                  # ctx read, gi_ctx write
        "K1_path_fwd": G + S + (S + S + S * S) + stash,          # gi_ctx, eps read; paths, means, chol, stash write
        "K5_elbo_fwd": 2 * S + S * S,                           # paths, means, chol read
        "K6_elbo_bwd": 2 * (2 * S + S * S),                     # same read + three cotangents written
        "K2_path_bwd": (2 * S + nt) + S + stash + NL * H + dg,  # cotangents, eps, stash (+h_prev) read; d_pre write
        "K3_grad_ctx": G + Cd,                                  # d_gi read, grad_ctx write
        "K4_wgrad": dg + (NL + 1) * H + Cd + S,                 # d_pre, h stash, ctx, paths read
    }[stage]


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half (idle clocks are low)
            out.update(sm_mhz=statistics.median(top), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_baseline(workload: str, n_steps_sample: int | None, iters: int, warm: int) -> dict:
    """Oracle (CPU port of the reference path: stepwise head + reference ELBO + autograd) timed on the
    host cores.  The ONLY place bench.py touches oracle/ (the checker, never the product)."""
    import torch

    from oracle import oracle_torch as O
    from viforsdes_b200.synthetic import WORKLOADS

    kind, B, T, dt = WORKLOADS[workload]
    Ts = T if n_steps_sample is None else min(T, n_steps_sample)
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core this process may run on
    try:
        torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    O.run_fwd_bwd(O.make_problem(kind, min(B, 16), 20, dt=dt))  # thread pools, allocator
    p = O.make_problem(kind, B, Ts, dt=dt)
    for _ in range(warm):
        O.run_fwd_bwd(p)
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        O.run_fwd_bwd(p)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": B * Ts / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{workload}: B={B}, first {Ts} of {T} grid steps, {iters} timed iteration(s), "
                      f"{sec * 1e3:.0f} ms each, torch {torch.__version__} CPU fp32, os.cpu_count()={os.cpu_count()}",
            "ms_per_iteration": sec * 1e3}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from viforsdes_b200.synthetic import WORKLOADS

    kind, B, T, dt = WORKLOADS[args.workload]
    cb = cpu_baseline(args.workload, REFERENCE_SAMPLE_STEPS, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_iteration"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "sde": kind, "batch_per_gpu": B, "n_steps": T, "dt": dt,
                   "note": "reference has no CPU implementation of this path (models/head.py:164-209 always launches "
                           "Triton); this is the oracle port of its PyTorch step math + ELBO + autograd"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="auto", choices=["auto", "generic", "fast", "tiled", "tc"],
                    help="recurrence kernel family (ad-hoc comparisons; production is auto)")
    ap.add_argument("--context-dtype", default="f32", choices=["f32", "bf16"],
                    help="dtype of context / grad_context in the device-resident leg (bf16 = the reference's AMP mode)")
    ap.add_argument("--no-graph", action="store_true", help="launch the iteration kernel by kernel in the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from viforsdes_b200 import _lib
    from viforsdes_b200.dist import allreduce_mean_, init_process_group
    from viforsdes_b200.runner import PathIteration
    from viforsdes_b200.session import HostSession
    from viforsdes_b200.synthetic import WORKLOADS, make_inputs

    os.environ["NCCL_DEBUG"] = os.environ.get("VISDE_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
    rank, local_rank, world = init_process_group()
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()

    kind, B, T, dt = WORKLOADS[args.workload]
    inp = make_inputs(kind, B, T, dt=dt, seed=rank)  # weak scaling: every rank its own B trajectories
    if world > 1:  # replicated parameters: all ranks use rank 0's weights
        ref = make_inputs(kind, 1, 1, dt=dt, seed=0)
        inp.w_ih, inp.w_hh, inp.b_ih, inp.b_hh, inp.out_w, inp.out_b = ref.w_ih, ref.w_hh, ref.b_ih, ref.b_hh, ref.out_w, ref.out_b
    variant = {"auto": _lib.VARIANT_AUTO, "generic": _lib.VARIANT_GENERIC, "fast": _lib.VARIANT_FAST,
               "tiled": _lib.VARIANT_TILED, "tc": _lib.VARIANT_TC}[args.variant]
    it = PathIteration(inp, dev, variant=variant,
                       context_dtype=torch.bfloat16 if args.context_dtype == "bf16" else torch.float32)
    S, Cd, H, NL = it.S, it.C, it.H, it.NL
    units = B * T
    flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)

    use_graph = not args.no_graph

    def step(graph: bool = False) -> None:
        if graph:
            it.replay()  # the whole iteration as one CUDA-graph launch (the NCCL exchange stays outside the graph)
        else:
            it.step()
        if world > 1:  # the path's one exchange step
            it.stage_elbo()               # the ELBO scalar rides in the tail slot of the gradient bucket:
            it.bucket.allreduce_mean_()   # ONE NCCL all-reduce (AVG) of 351 KB per iteration

    K, W = args.steps, args.warmup
    with ClockSampler(local_rank) as clk:
        for _ in range(W):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if use_graph:
            it.capture()
            for _ in range(2):
                step(graph=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
        # ---- timed region: K iterations, CUDA events per step on the launching stream ----
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        t_wall = time.perf_counter()
        for i in range(K):
            flush.zero_()  # L2 flush between timed iterations (outside the event bracket)
            ev[i][0].record()
            step(graph=use_graph)
            ev[i][1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_wall = time.perf_counter() - t_wall
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # ---- per-stage breakdown: the same K iterations kernel by kernel with the library's stage timer ----
        _lib.check(lib.visde_profile_begin(K * 8 + 8))
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.zero_()
            ev2[i][0].record()
            step()
            ev2[i][1].record()
        torch.cuda.synchronize()
        ms = (C.c_double * len(_lib.STAGES))()
        cnt = (C.c_int * len(_lib.STAGES))()
        _lib.check(lib.visde_profile_end(ms, cnt))
        eager_ms_per_step = sum(a.elapsed_time(b) for a, b in ev2) / K
        # ---- informational: a complete data-parallel head iteration = the timed step + the fused unscale / clip / AdamW / EMA
        # over the flat parameter and gradient buffers (lr = 0 keeps the weights: same kernels and traffic, same inputs later)
        opt = it.make_optimizer(lr=0.0, max_norm=1.0, ema_decay=0.999)
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.zero_()
            ev3[i][0].record()
            step(graph=use_graph)
            opt.step()
            ev3[i][1].record()
        torch.cuda.synchronize()
        train_ms_per_step = sum(a.elapsed_time(b) for a, b in ev3) / K
        tt = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_per_step = tt.item() / K
        value = units * world / (ms_per_step * 1e-3)

        # ---- e2e: host buffers through the C-ABI session (H2D + D2H inside the timed region) ----
        e2e = None
        if not args.no_e2e:
            sess = HostSession.from_inputs(inp)
            for _ in range(max(2, W // 2)):
                sess.step()
            # (a) synchronous call per step: latency of one iteration through host buffers
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                sess.step()
            t_sync = time.perf_counter() - t0
            # (b) the pipelined form of the same call (visde_session_submit / _wait): every step still copies
            # its own inputs from pinned host memory and reads its own results back; the H2D of step i+1
            # overlaps the kernels of step i (two device input sets)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            sess.submit()
            for _ in range(K - 1):
                sess.submit()
                sess.wait()
            sess.wait()
            te = torch.tensor([time.perf_counter() - t0, t_sync], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            t_pipe, t_sync = te[0].item(), te[1].item()
            e2e = {"value": units * world * K / t_pipe, "unit": UNIT, "h2d_bytes_per_step": sess.h2d_bytes,
                   "d2h_bytes_per_step": sess.d2h_bytes, "ms_per_step": t_pipe / K * 1e3,
                   "sync_ms_per_step": t_sync / K * 1e3, "sync_value": units * world * K / t_sync,
                   "api": "visde_session_submit/_wait (C ABI, pinned host buffers, 2 iterations in flight: H2D of step "
                          "i+1 overlaps the kernels of step i); sync_* = visde_session_step, one blocking call per "
                          "step; grad_context stays on device for the encoder backward"}
            sess.close()
    clocks = clk.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    stages = {n: {"ms_per_step": ms[i] / K, "launches_per_step": cnt[i] / K,
                  "algorithmic_bytes_per_unit": stage_bytes_per_unit(n, S, Cd, H, NL)}
              for i, n in enumerate(_lib.STAGES)}
    for n, s in stages.items():
        s["achieved_gbs"] = (s["algorithmic_bytes_per_unit"] * units / (s["ms_per_step"] * 1e-3) / 1e9) if s["ms_per_step"] > 0 else None
    dom = max(stages, key=lambda n: stages[n]["ms_per_step"])
    d = stages[dom]
    traffic = None
    tj = ROOT / "profiles" / "ncu_traffic.json"
    if tj.exists():
        traffic = json.loads(tj.read_text()).get(args.workload, {}).get(dom)
    dur = d["ms_per_step"] / max(1.0, d["launches_per_step"]) * 1e-3
    roofline = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": (d["achieved_gbs"] / hbm_peak) if d["achieved_gbs"] else None, "traffic": traffic,
                "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback 6.65 TB/s",
                "kernel_ms": dur * 1e3, "share_of_step": d["ms_per_step"] / eager_ms_per_step,
                "timing": "CUDA events of the library's stage timer around this stage, same K iterations launched kernel by "
                          "kernel right after the timed region (the timed region itself is one graph replay per iteration)",
                "note": ("B=128 trajectories on 148 SMs with T serial steps: latency-bound, see DESIGN.md §5; "
                         "whole-path figures in step_roofline") if B <= 148 else
                        ("large-batch family: gate GEMMs on tcgen05, HBM-bound by the stash / d_pre streams "
                         "(DESIGN.md §5); whole-path figures in step_roofline")}
    fl = path_flops_per_unit(S, Cd, H, NL)
    step_roofline = {"path_bytes_per_unit": path_bytes_per_unit(S, Cd), "path_flops_per_unit": fl,
                     "achieved_gbs": path_bytes_per_unit(S, Cd) * units / (ms_per_step * 1e-3) / 1e9,
                     "achieved_tflops": fl * units / (ms_per_step * 1e-3) / 1e12,
                     "fp32_simt_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12}

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args.workload, None, 1, 0)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "sde": kind, "batch_per_gpu": B, "n_steps": T, "dt": dt, "state_dim": S,
                   "context_dim": Cd, "hidden_dim": H, "num_layers": NL, "parallelism": f"dp{world}", "variant": args.variant, "context_dtype": args.context_dtype,
                   "launch": "one CUDA-graph replay per iteration" if use_graph else "kernel by kernel",
                   "l2": "512 MB flush write between timed steps; per-step working set ~0.9 GB > 126 MB L2"},
        "e2e": e2e, "gpu_launches": int(sum(cnt)),
        "roofline": roofline, "step_roofline": step_roofline, "stages": stages, "cpu_baseline": cb,
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"]},
        "wall_ms_per_step_incl_flush": t_wall / K * 1e3, "eager_ms_per_step": eager_ms_per_step,
        "head_train_ms_per_step": train_ms_per_step,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
