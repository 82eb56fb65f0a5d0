"""Two ranks on two GPUs over NCCL: the averaged shard gradients of the data-parallel iteration
equal the single-GPU gradients on the concatenated batch with the same injected theta / eps; the ELBO scalar rides
in the bucket's tail slot; the EMA shadow average and the full-model gradient bucket go through the same collective.
Needs >= 2 CUDA devices (`gpurun --gpus 2`); skipped otherwise."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, kind: str, B: int, T: int, capture: bool) -> None:
    import dataclasses

    import torch.distributed as dist

    from oracle import oracle_torch as O
    from tests._util import normwise
    from tests.test_gpu_config5 import _inputs_from_problem
    from viforsdes_b200.dist import EmaSync, allreduce_grads_, init_process_group, shard_range
    from viforsdes_b200.runner import PathIteration

    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    init_process_group("nccl")
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    try:
        p = O.make_problem(kind, B, T, context_dim=128, hidden_dim=64, num_layers=2, state_dim=10 if kind == "l96" else None)
        lo, hi = shard_range(B, rank, world)
        shard = dataclasses.replace(p, x0=p.x0[lo:hi], context=p.context[lo:hi], theta=p.theta[lo:hi], eps=p.eps[lo:hi])
        it = PathIteration(_inputs_from_problem(shard), dev)

        def exchange():
            it.stage_elbo()
            it.bucket.allreduce_mean_()

        it.step()
        exchange()
        if capture:  # the iteration and its NCCL all-reduce replayed as ONE CUDA graph (what bench.py times)
            it.capture(post=exchange)
            it.bucket.flat.zero_()
            it.replay()
        torch.cuda.synchronize()
        full = PathIteration(_inputs_from_problem(p), dev)  # the same batch on one GPU
        full.step()
        full.stage_elbo()
        for a, b, nm in zip(it.head_weight_grads(), full.head_weight_grads(), range(100)):
            assert normwise(a, b) < 2e-5, f"rank {rank}: averaged shard gradient {nm} differs from the single-GPU gradient: {normwise(a, b)}"
        assert abs(it.bucket.extra.item() - full.bucket.extra.item()) <= 2e-5 * abs(full.bucket.extra.item()), "ELBO scalar"
        # per-trajectory gradients of the shard are the rows of the full-batch ones (scaled by the batch-mean factor)
        r_s, r_f = it.results()["grads"], full.results()["grads"]
        for nm in ("x0", "theta", "context"):
            assert normwise(r_s[nm] / world, r_f[nm][lo:hi]) < 2e-5, nm
        # the full-model collectives on the same communicator: arbitrary module gradients and the EMA shadow
        lin = torch.nn.Linear(7, 5).to(dev)
        for q in lin.parameters():
            q.grad = torch.full_like(q, float(rank + 1))
        allreduce_grads_(lin.parameters())
        assert all(torch.allclose(q.grad, torch.full_like(q, (world + 1) / 2)) for q in lin.parameters())
        shadow = torch.full((1000,), float(rank), device=dev)
        EmaSync([shadow], every=1).step()
        assert torch.allclose(shadow, torch.full_like(shadow, (world - 1) / 2))
        torch.cuda.synchronize()
        dist.barrier()
        if hasattr(it, "graph"):
            del it.graph  # the captured iteration holds NCCL kernels: release it before the communicator goes away
        torch.cuda.synchronize()
    finally:
        import threading

        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(20.0)  # a blocked teardown must not hang the test run
        if th.is_alive():
            os._exit(0)


@pytest.mark.timeout(180)
@pytest.mark.parametrize("kind,B,T,capture", [("lv", 256, 30, False), ("l96", 256, 20, True)])
def test_two_rank_nccl_gradients_match_single_gpu(kind, B, T, capture):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    mp.spawn(_worker, args=(2, _free_port(), kind, B, T, capture), nprocs=2, join=True)
