#!/usr/bin/env python
"""Benchmark of the variational path-sampling hot path (BASELINE.json metric: trajectory-steps/s
for fwd + bwd ELBO at 1/2/4/8 B200, ms per ELBO iteration).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

Default workload = the configuration the metric is quoted on, BASELINE.json configs[4]: the 10-D stochastic
Lorenz-96 user-defined SDE, 8192 trajectories, T = 100, STRONG-scaled over the N ranks (8192 / N trajectories per
rank, inference/trainer.py:169 + training_context.py:59-68), one NCCL all-reduce of the head's gradient bucket per
iteration.  A step = one pass of the hot path over the batch:
  K0 ctx GEMM -> K1 path fwd -> [user drift/diffusion in PyTorch] -> K5 ELBO fwd -> K6 ELBO bwd -> [their VJP] ->
  K2 path bwd -> K3 grad_ctx -> K4 wgrads  (+ for N > 1 the all-reduce, captured in the same CUDA graph).
`value`   : inputs resident in HBM, CUDA-event time per step on the launching stream, max over ranks.
`e2e`     : the same iteration through the host-buffer C-ABI session (visde_session_submit/_wait): pinned HOST
            buffers in (bf16 context = the reference's autocast mode), H2D + kernels + D2H inside the timed region.
`roofline`: dominant kernel (by device time, from the library's stage profiler) against the measured HBM peak, plus
            the whole-path contract roofline of SURVEY.md §8(d) (`contract_*`, `bound`, `traffic_ratio`).
`cpu_baseline` / `--impl reference`: the CPU oracle (stepwise PyTorch port of the reference path; the reference has
            no CPU implementation of it) on the host cores, on a stated bounded sample of the same workload.
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
DEFAULT_WORKLOAD = "l96_b8192_t100"  # BASELINE.json configs[4]: the configuration the metric is quoted on
CPU_SAMPLE_TRAJECTORIES = 1024       # CPU arms: the first 1024 trajectories of the batch, ALL grid steps
EXTRA_WORKLOAD = "ou_b16384_t100"    # configs[2] roofline characterisation, reported under `extra`
FULL_MODEL_PARAMS = 8_278_831        # encoder 8.19 M + head 87.7 k + posterior 6: the DDP bucket
CONTEXT_DIM, HIDDEN_DIM, NUM_LAYERS = 256, 64, 2


def path_bytes_per_unit(S: int, Cd: int) -> int:
    """operator-contract HBM bytes per trajectory-step, fp32 I/O, stash excluded."""
    return 4 * (3 * Cd + 13 * S + 5 * S * S)


def path_flops_per_unit(S: int, Cd: int, H: int, NL: int) -> int:
    G, nt = 3 * H, S * (S + 1) // 2
    mac = (S + Cd) * G + H * G + (NL - 1) * 2 * H * G + H * (S + nt) + nt
    return 6 * mac


def stage_bytes_per_unit(stage: str, S: int, Cd: int) -> int:
    """Per-kernel ALGORITHMIC bytes per trajectory-step: the operator-contract tensors that kernel touches (they
    sum to path_bytes_per_unit).  The stash, gi_ctx and d_pre streams are implementation choices and are NOT
    counted here; they show up in `traffic` (ncu dram bytes)."""
    return 4 * {
        "K0_ctx_gemm": Cd,                       # ctx read
        "K1_path_fwd": 3 * S + S * S,            # eps read; paths, means, chol write
        "K5_elbo_fwd": 2 * S + S * S,            # paths, means, chol read
        "K6_elbo_bwd": 4 * S + 2 * S * S,        # same read + gP, gM, gL write
        "K2_path_bwd": 4 * S + S * S,            # gP, gM, gL, eps, paths read
        "K3_grad_ctx": Cd,                       # grad_ctx write
        "K4_wgrad": Cd,                          # ctx read
    }[stage]


def workload_config(name: str) -> dict:
    """The `config` object both arms print (identical by construction)."""
    from viforsdes_b200.synthetic import WORKLOADS

    kind, B, T, dt = WORKLOADS[name]
    S = {"ou": 1, "lv": 2, "l96": 10}[kind]
    return {"workload": name, "sde": kind, "global_batch": B, "n_steps": T, "dt": dt, "state_dim": S,
            "context_dim": CONTEXT_DIM, "hidden_dim": HIDDEN_DIM, "num_layers": NUM_LAYERS}


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


def cpu_baseline(workload: str, iters: int, warm: int) -> dict:
    """Oracle (CPU port of the reference path: stepwise head + reference ELBO + autograd) timed on the host cores on
    a bounded sample: the first CPU_SAMPLE_TRAJECTORIES trajectories of the workload's batch over ALL its grid steps
    (trajectories are independent, so trajectory-steps/s does not depend on the sample size beyond cache effects).
    The ONLY place bench.py touches oracle/ (the checker, never the product)."""
    import torch

    from oracle import oracle_torch as O
    from viforsdes_b200.synthetic import WORKLOADS

    kind, B, T, dt = WORKLOADS[workload]
    Bs = min(B, CPU_SAMPLE_TRAJECTORIES)
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core this process may run on
    try:
        torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    O.run_fwd_bwd(O.make_problem(kind, min(B, 16), 20, dt=dt))  # thread pools, allocator
    p = O.make_problem(kind, Bs, T, dt=dt, context_dim=CONTEXT_DIM, hidden_dim=HIDDEN_DIM, num_layers=NUM_LAYERS)
    for _ in range(warm):
        O.run_fwd_bwd(p)
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        O.run_fwd_bwd(p)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": Bs * T / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{workload}: first {Bs} of {B} trajectories, all {T} grid steps, {iters} timed iteration(s), "
                      f"{sec * 1e3:.0f} ms each, torch {torch.__version__} CPU fp32, os.cpu_count()={os.cpu_count()}; "
                      "the reference has no CPU implementation of this path (models/head.py:164-209 always launches "
                      "Triton): this is the oracle port of its PyTorch step math + ELBO + autograd",
            "ms_per_iteration": sec * 1e3}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cb = cpu_baseline(args.workload, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_iteration"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def timed_steps(fn, K: int, flush, torch):
    """K calls of fn, each bracketed by CUDA events on the current stream, L2 flushed in between; total ms."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.zero_()  # L2 flush between timed iterations (outside the event bracket)
        ev[i][0].record()
        fn()
        ev[i][1].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="auto", choices=["auto", "generic", "fast", "tiled", "tc"],
                    help="recurrence kernel family (ad-hoc comparisons; production is auto)")
    ap.add_argument("--context-dtype", default="f32", choices=["f32", "bf16"],
                    help="dtype of context / grad_context in the device-resident leg (bf16 = the reference's AMP mode)")
    ap.add_argument("--e2e-context-dtype", default="bf16", choices=["f32", "bf16"],
                    help="dtype of the HOST context of the e2e leg (bf16 = the reference's autocast encoder output)")
    ap.add_argument("--e2e-noise", default="device", choices=["device", "host"],
                    help="e2e leg: eps drawn on the device per iteration (what the reference's sampler does) or copied from host")
    ap.add_argument("--no-graph", action="store_true", help="launch the iteration kernel by kernel in the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` legs (config-3 roofline point, comm)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from viforsdes_b200 import _lib
    from viforsdes_b200.dist import EmaSync, init_process_group
    from viforsdes_b200.runner import PathIteration
    from viforsdes_b200.session import HostSession
    from viforsdes_b200.synthetic import WORKLOADS, make_inputs

    rank, local_rank, world = init_process_group()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()

    kind, B_global, T, dt = WORKLOADS[args.workload]
    # strong scaling whenever the global batch shards into >= 128 trajectories per rank (config 5: 8192 / N);
    # the B = 128 latency configurations replicate the batch per rank (weak)
    strong = B_global % world == 0 and B_global // world >= 128
    B = B_global // world if strong else B_global
    inp = make_inputs(kind, B, T, dt=dt, seed=rank, context_dim=CONTEXT_DIM, hidden_dim=HIDDEN_DIM, num_layers=NUM_LAYERS)
    if world > 1:  # replicated parameters: all ranks use rank 0's weights
        ref = make_inputs(kind, 1, 1, dt=dt, seed=0, context_dim=CONTEXT_DIM, hidden_dim=HIDDEN_DIM, num_layers=NUM_LAYERS)
        inp.w_ih, inp.w_hh, inp.b_ih, inp.b_hh, inp.out_w, inp.out_b = ref.w_ih, ref.w_hh, ref.b_ih, ref.b_hh, ref.out_w, ref.out_b
    variant = {"auto": _lib.VARIANT_AUTO, "generic": _lib.VARIANT_GENERIC, "fast": _lib.VARIANT_FAST,
               "tiled": _lib.VARIANT_TILED, "tc": _lib.VARIANT_TC}[args.variant]
    it = PathIteration(inp, dev, variant=variant,
                       context_dtype=torch.bfloat16 if args.context_dtype == "bf16" else torch.float32)
    S, Cd, H, NL = it.S, it.C, it.H, it.NL
    units_local = B * T
    units = units_local * world
    flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
    use_graph = not args.no_graph
    K, W = args.steps, args.warmup

    def exchange() -> None:
        """The path's one exchange step: the ELBO scalar rides in the tail slot of the head's flat
        gradient bucket, so it is ONE NCCL all-reduce (AVG) per iteration."""
        if world > 1:
            it.stage_elbo()
            it.bucket.allreduce_mean_()

    def eager_step() -> None:
        it.step()
        exchange()

    comm_in_graph = False
    with ClockSampler(local_rank) as clk:
        for _ in range(W):
            eager_step()  # also initialises the NCCL communicator before any capture
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        graph_step = eager_step
        if use_graph:
            # the whole iteration INCLUDING the all-reduce as one CUDA graph (NCCL is capturable): no launch gap
            # between the last kernel and the collective
            try:
                it.capture(post=exchange if world > 1 else None)
                comm_in_graph = world > 1
                graph_step = it.replay
            except Exception as e:  # capture of the collective unavailable: kernels in the graph, NCCL behind it
                if world == 1:
                    raise
                print(f"[bench] capturing the all-reduce failed ({type(e).__name__}: {e}); launching it after the replay",
                      file=sys.stderr)
                torch.cuda.synchronize()
                it.capture()

                def graph_step() -> None:
                    it.replay()
                    exchange()
            for _ in range(2):
                graph_step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
        # ---- timed region: K iterations, CUDA events per step on the launching stream ----
        t_wall = time.perf_counter()
        dev_ms = timed_steps(graph_step, K, flush, torch)
        if world > 1:
            dist.barrier()
        t_wall = time.perf_counter() - t_wall
        # ---- per-stage breakdown: the same K iterations kernel by kernel with the library's stage timer ----
        _lib.check(lib.visde_profile_begin(K * 8 + 8))
        eager_ms_per_step = timed_steps(eager_step, K, flush, torch) / K
        ms = (C.c_double * len(_lib.STAGES))()
        cnt = (C.c_int * len(_lib.STAGES))()
        _lib.check(lib.visde_profile_end(ms, cnt))
        # ---- informational: a complete data-parallel head iteration = the timed step + the fused unscale / clip / AdamW /
        # EMA over the flat parameter and gradient buffers (lr = 0 keeps the weights: same kernels and traffic)
        opt = it.make_optimizer(lr=0.0, max_norm=1.0, ema_decay=0.999)

        def train_step() -> None:
            graph_step()
            opt.step()

        train_ms_per_step = timed_steps(train_step, K, flush, torch) / K
        tt = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_per_step = tt.item() / K
        value = units / (ms_per_step * 1e-3)

        extra: dict = {}
        # ---- extra.comm: the collectives of the FULL model on hardware (README.md:97, training_context.py:91):
        # the 33 MB DDP gradient bucket and the EMA-shadow sync, alone and overlapped with the iteration on a side stream
        if world > 1 and not args.no_extra:
            full = torch.zeros(FULL_MODEL_PARAMS, device=dev)
            shadow = torch.zeros(FULL_MODEL_PARAMS, device=dev)
            ema = EmaSync([shadow], every=1)
            side = torch.cuda.Stream(dev)

            def time_alone(fn) -> float:
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                dist.barrier()
                t = torch.tensor([timed_steps(fn, 10, flush, torch) / 10], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return t.item()

            from viforsdes_b200.dist import allreduce_mean_
            ar_ms = time_alone(lambda: allreduce_mean_(full))
            ema_ms = time_alone(ema.sync)

            def step_with_full_comm() -> None:
                # encoder-gradient bucket + EMA shadow average on a side stream behind the previous iteration's
                # gradients; the path kernels of this iteration run concurrently on the main stream
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    allreduce_mean_(full)
                    ema.sync()
                graph_step()
                torch.cuda.current_stream(dev).wait_stream(side)

            for _ in range(2):
                step_with_full_comm()
            torch.cuda.synchronize()
            dist.barrier()
            t = torch.tensor([timed_steps(step_with_full_comm, K, flush, torch) / K], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            extra["comm"] = {
                "full_model_params": FULL_MODEL_PARAMS, "bucket_mb": FULL_MODEL_PARAMS * 4 / 1e6,
                "full_model_allreduce_ms": ar_ms, "ema_sync_ms": ema_ms,
                "allreduce_busbw_gbs": FULL_MODEL_PARAMS * 4 * 2 * (world - 1) / world / (ar_ms * 1e-3) / 1e9,
                "ms_per_step_with_full_model_allreduce_and_ema_sync_overlapped": t.item(),
                "ms_per_step_path_only": ms_per_step,
                "head_bucket_bytes": it.bucket.flat.numel() * 4, "head_allreduce_in_graph": comm_in_graph,
                "note": "33 MB synthetic bucket = the reference model's 8.28 M fp32 gradients (DDP, training_context.py:91) "
                        "and its EMA shadow (README.md:97), averaged with NCCL on a side stream while the path kernels run",
            }
            del full, shadow

        # ---- e2e: host buffers through the C-ABI session (H2D + D2H inside the timed region) ----
        e2e = None
        if not args.no_e2e:
            e2e_dt = torch.bfloat16 if args.e2e_context_dtype == "bf16" else torch.float32
            sess = HostSession.from_inputs(inp, context_dtype=e2e_dt,
                                           device_noise_seed=(1234 + rank) if args.e2e_noise == "device" else None)
            for _ in range(max(6, W // 2)):  # past the session's eager warm-up: user-SDE hooks are CUDA graphs from the 4th call
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
            e2e = {"value": units * K / t_pipe, "unit": UNIT, "h2d_bytes_per_step": sess.h2d_bytes,
                   "d2h_bytes_per_step": sess.d2h_bytes, "ms_per_step": t_pipe / K * 1e3,
                   "sync_ms_per_step": t_sync / K * 1e3, "sync_value": units * K / t_sync,
                   "host_context_dtype": args.e2e_context_dtype,
                   "noise": ("drawn on the device inside the timed region, a fresh Philox stream per iteration (the reference's "
                             "sampler draws torch.randn on the device, inference/diffusion_path_sampler.py:57)"
                             if args.e2e_noise == "device" else "copied from pinned host memory every step"),
                   "api": "visde_session_submit/_wait (C ABI, pinned host buffers, 2 iterations in flight: H2D of step "
                          "i+1 overlaps the kernels of step i); the user SDE's drift / diffusion + VJP run in PyTorch "
                          "through the session's visde_user_sde hooks; sync_* = visde_session_step, one blocking call "
                          "per step; bytes are per rank; grad_context stays on device for the encoder backward"}
            sess.close()
            del sess
    clocks = clk.summary()

    # ---- extra.roofline_point: BASELINE configs[2] (OU batch sweep) at B = 16 384 on this rank's GPU, so that the
    # large-batch tensor-core family's roofline fraction is visible in the default line
    if rank == 0 and world == 1 and not args.no_extra and args.workload != EXTRA_WORKLOAD:
        try:
            del it, opt, inp
            torch.cuda.empty_cache()
            extra["roofline_point"] = roofline_point(EXTRA_WORKLOAD, dev, flush, max(5, K // 2))
        except Exception as e:  # informational leg: never fail the headline
            extra["roofline_point"] = {"error": f"{type(e).__name__}: {e}"}

    if world > 1:
        shutdown_ranks(rank, locals().get("it"))
    if rank != 0:
        return

    # ---- roofline --------------------------------------------------------------------------------
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    stages = {n: {"ms_per_step": ms[i] / K, "launches_per_step": cnt[i] / K,
                  "algorithmic_bytes_per_unit": stage_bytes_per_unit(n, S, Cd)}
              for i, n in enumerate(_lib.STAGES)}
    for n, s in stages.items():
        s["achieved_gbs"] = (s["algorithmic_bytes_per_unit"] * units_local / (s["ms_per_step"] * 1e-3) / 1e9) if s["ms_per_step"] > 0 else None
    stage_sum = sum(s["ms_per_step"] for s in stages.values())
    dom = max(stages, key=lambda n: stages[n]["ms_per_step"])
    d = stages[dom]
    traffic = traffic_path = None
    tj = ROOT / "profiles" / "ncu_traffic.json"
    if tj.exists():
        tw = json.loads(tj.read_text()).get(args.workload, {})
        traffic = tw.get(dom)
        if all(n in tw for n in _lib.STAGES):
            traffic_path = sum(tw[n] for n in _lib.STAGES)
    family = _lib.FAMILY_NAMES[lib.visde_recurrence_family(C.byref(it_dims(B, T, S, Cd, inp_P(kind), H, NL, variant)), 1)]
    roofline = {"kernel": dom, "achieved": d["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": (d["achieved_gbs"] / hbm_peak) if d["achieved_gbs"] else None, "traffic": traffic,
                "bound": "hbm",
                "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback 6.65 TB/s",
                "kernel_ms": d["ms_per_step"] / max(1.0, d["launches_per_step"]),
                "share_of_step": d["ms_per_step"] / eager_ms_per_step,
                "timing": "CUDA events of the library's stage timer around this stage, same K iterations launched kernel by "
                          "kernel right after the timed region (the timed region itself is one graph replay per iteration)",
                "algorithmic_bytes": "operator-contract tensors this kernel touches (SURVEY.md §8d; stash / gi_ctx / d_pre "
                                     "streams excluded: they are in `traffic`)"}
    if dom in ("K1_path_fwd", "K2_path_bwd"):
        # the recurrence stages are not HBM-bound: their roof is the unit that executes the gate products.
        # Algorithmic FLOPs of the stage = 2 x the recurrent MACs (K1: products with W; K2: the transposed products)
        G, nt = 3 * H, S * (S + 1) // 2
        rec_flops = 2 * (S * G + H * G + (NL - 1) * 2 * H * G + H * (S + nt) + nt)
        tfl = rec_flops * units_local / (d["ms_per_step"] * 1e-3) / 1e12
        if family == "tc":
            unit_peak = (peaks.get("bf16_tflops_sustained") or 1360.5) / 3.0
            bound, unit_name = "tensor", "tcgen05 kind::f16, fp16 hi/lo 3-pass split: a third of the measured dense 16-bit rate"
        else:
            unit_peak = 148 * 128 * 2 * 1.965e9 / 1e12
            bound, unit_name = "fp32_simt", "FFMA2 on 148 SMs x 128 lanes x 1.965 GHz"
        roofline.update({"bound": bound, "achieved": tfl, "peak": unit_peak, "unit": "TFLOP/s", "frac": tfl / unit_peak,
                         "peak_source": unit_name, "algorithmic_flops_per_unit": rec_flops,
                         "hbm_achieved_gbs": d["achieved_gbs"], "hbm_frac": (d["achieved_gbs"] / hbm_peak) if d["achieved_gbs"] else None,
                         "note": "latency-bound by construction: T serial steps per trajectory tile, "
                                 f"{(B + 127) // 128 if family == 'tc' else min(B, 148)} of 148 SMs busy (DESIGN.md)"})
    cr = contract_roofline(S, Cd, H, NL, units_local, ms_per_step, family, hbm_peak, peaks, traffic_path)
    cr["path_bound"] = cr.pop("bound")
    roofline.update(cr)

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(args.workload, 3, 1)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    cfg = workload_config(args.workload)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "run": {"batch_per_gpu": B, "parallelism": f"dp{world}", "variant": args.variant, "recurrence_family": family,
                "context_dtype": args.context_dtype,
                "launch": ("one CUDA-graph replay per iteration" + (" incl. the NCCL all-reduce" if comm_in_graph else ""))
                if use_graph else "kernel by kernel",
                "l2": "512 MB flush write between timed steps; per-step working set > 126 MB L2"},
        "e2e": e2e, "gpu_launches": int(sum(cnt)),
        "roofline": roofline, "stages": stages, "stage_sum_ms": stage_sum,
        "outside_stage_ms_per_step": eager_ms_per_step - stage_sum,  # PyTorch user SDE + VJP, gradient adds, (NCCL)
        "cpu_baseline": cb,
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"]},
        "wall_ms_per_step_incl_flush": t_wall / K * 1e3, "eager_ms_per_step": eager_ms_per_step,
        "head_train_ms_per_step": train_ms_per_step, "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        os._exit(0)  # the process group was already torn down in shutdown_ranks


def shutdown_ranks(rank: int, it) -> None:
    """Orderly end of a multi-rank run.  The captured iteration holds NCCL kernels: the graph is released and the device
    drained BEFORE the communicator goes away (destroying a process group under a live graph can block forever), and a
    watchdog ends the process if the teardown still does not return.  Non-zero ranks exit here; rank 0 goes on to print."""
    import threading

    import torch
    import torch.distributed as dist

    torch.cuda.synchronize()
    dist.barrier()
    if it is not None and hasattr(it, "graph"):
        del it.graph
    torch.cuda.synchronize()
    def teardown() -> None:
        try:
            dist.destroy_process_group()
        except Exception:
            pass

    th = threading.Thread(target=teardown, daemon=True)
    th.start()
    th.join(20.0)  # a teardown that does not return must not keep the ranks (and the driver's clock) alive
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)


def inp_P(kind: str) -> int:
    return {"ou": 3, "lv": 3, "l96": 2}[kind]


def it_dims(B, T, S, Cd, P, H, NL, variant):
    from viforsdes_b200 import _lib

    return _lib.Dims(B, T, S, Cd, P, H, NL, variant)


def contract_roofline(S, Cd, H, NL, units, ms_per_step, family, hbm_peak, peaks, traffic_path) -> dict:
    """rate / min(BW / bytes, PEAK_unit / FLOPs) with the unit that executed the gate products named.
    FP32 SIMT families: 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s.  Tensor-core family: every product is a
    3-pass split (fp16 hi/lo in the recurrence at the measured dense 16-bit rate; 3xTF32 at half that rate for the
    context GEMMs, 57 % of the MACs), so the usable rate on ALGORITHMIC flops is a third of the pipe's."""
    by, fl = path_bytes_per_unit(S, Cd), path_flops_per_unit(S, Cd, H, NL)
    rate = units / (ms_per_step * 1e-3)
    simt = 148 * 128 * 2 * 1.965e9
    t16 = (peaks.get("bf16_tflops_sustained") or 1360.5) * 1e12
    ctx_share = 6 * Cd * 3 * H / fl
    hbm_rate = hbm_peak * 1e9 / by
    if family == "tc":
        unit_name = "tensor (tcgen05 kind::f16 hi/lo 3-pass recurrence + 3xTF32 context GEMMs)"
        unit_rate = 1.0 / (fl * ctx_share * 3 / (t16 / 2) + fl * (1 - ctx_share) * 3 / t16)
        bound_unit = "tensor"
    else:
        unit_name = "fp32_simt (FFMA2 gate products; context GEMMs on tcgen05 3xTF32)"
        unit_rate = 1.0 / (fl * ctx_share * 3 / (t16 / 2) + fl * (1 - ctx_share) / simt)
        bound_unit = "fp32_simt"
    roof = min(hbm_rate, unit_rate)
    out = {"bound": "hbm" if hbm_rate <= unit_rate else bound_unit, "executing_unit": unit_name,
           "contract_bytes_per_unit": by, "contract_flops_per_unit": fl,
           "contract_hbm_roof_units_per_s": hbm_rate, "contract_unit_roof_units_per_s": unit_rate,
           "contract_frac": rate / roof, "contract_hbm_frac": rate / hbm_rate,
           "achieved_contract_gbs": by * rate / 1e9, "achieved_tflops": fl * rate / 1e12,
           "traffic_ratio": (traffic_path / (by * units)) if traffic_path else None}
    return out


def roofline_point(workload: str, dev, flush, K: int) -> dict:
    """One device-resident measurement of another BASELINE configuration (no e2e, no CPU arm), for `extra`."""
    import torch

    from viforsdes_b200 import _lib
    from viforsdes_b200.runner import PathIteration
    from viforsdes_b200.synthetic import WORKLOADS, make_inputs

    lib = _lib.load()
    kind, B, T, dt = WORKLOADS[workload]
    inp = make_inputs(kind, B, T, dt=dt, seed=0, context_dim=CONTEXT_DIM, hidden_dim=HIDDEN_DIM, num_layers=NUM_LAYERS)
    it = PathIteration(inp, dev)
    for _ in range(3):
        it.step()
    it.capture()
    for _ in range(2):
        it.replay()
    torch.cuda.synchronize()
    ms_per_step = timed_steps(it.replay, K, flush, torch) / K
    _lib.check(lib.visde_profile_begin(K * 8 + 8))
    timed_steps(it.step, K, flush, torch)
    ms = (C.c_double * len(_lib.STAGES))()
    cnt = (C.c_int * len(_lib.STAGES))()
    _lib.check(lib.visde_profile_end(ms, cnt))
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    family = _lib.FAMILY_NAMES[lib.visde_recurrence_family(C.byref(it.dims), 1)]
    traffic_path = None
    tj = ROOT / "profiles" / "ncu_traffic.json"
    if tj.exists():
        tw = json.loads(tj.read_text()).get(workload, {})
        if all(n in tw for n in _lib.STAGES):
            traffic_path = sum(tw[n] for n in _lib.STAGES)
    out = {"workload": workload, "config": workload_config(workload), "ms_per_step": ms_per_step,
           "value": B * T / (ms_per_step * 1e-3), "unit": UNIT, "recurrence_family": family,
           "stages_ms": {n: ms[i] / K for i, n in enumerate(_lib.STAGES)}}
    out.update(contract_roofline(it.S, it.C, it.H, it.NL, B * T, ms_per_step, family, hbm_peak, peaks, traffic_path))
    return out


if __name__ == "__main__":
    main()
