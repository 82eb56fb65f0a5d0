"""Data-parallel plumbing for the path: trajectories are sharded across ranks
with no data-path collective; the only exchange per iteration is the average of the head's
weight gradients (one flat fp32 bucket, ~351 KB at H=64x2) and of the scalar ELBO, plus a
periodic average of the EMA shadow (claimed by the reference's README.md:97 but absent from
inference/exponential_moving_average.py).  One process per GPU, ``torch.distributed`` (NCCL on
NVLink; gloo in the CPU tests); reads RANK / LOCAL_RANK / WORLD_SIZE like
inference/training_context.py:59-68."""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def env_rank() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend: str | None = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(total: int, rank: int, world: int, allow_uneven: bool = False) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` trajectories for `rank`.  The gradient exchange averages per-rank batch
    MEANS with equal weight (`allreduce_mean_`), which is the global-batch mean only for equal shards: an uneven split
    is refused unless the caller says it rescales (`shard_weight`) before the all-reduce."""
    base, rem = divmod(total, world)
    if rem and not allow_uneven:
        raise ValueError(f"{total} trajectories do not split evenly over {world} ranks; pass allow_uneven=True and "
                         "scale each rank's gradients / ELBO by shard_weight() before the all-reduce")
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_weight(total: int, rank: int, world: int) -> float:
    """local_B * world / total: the factor that turns an equal-weight average of per-rank batch means into the
    global-batch mean when shards are uneven (1.0 for an even split)."""
    lo, hi = shard_range(total, rank, world, allow_uneven=True)
    return (hi - lo) * world / total


class FlatBucket:
    """One contiguous fp32 buffer whose slices ARE the gradient tensors (no pack / unpack copies)."""

    def __init__(self, shapes: Sequence[Sequence[int]], device: torch.device | str, extra: int = 0) -> None:
        """`extra` scalars ride at the tail of the bucket (the ELBO value: one collective per iteration instead
        of two)."""
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n + 3) // 4 * 4  # keep every view 16-byte aligned
        self.flat = torch.zeros(total + extra, device=device, dtype=torch.float32)
        self.views: List[Tensor] = [self.flat[o:o + n].view(*s) for o, n, s in zip(offs, sizes, shapes)]
        self.extra: Tensor = self.flat[total:total + extra]

    def allreduce_mean_(self, group=None) -> None:
        allreduce_mean_(self.flat, group)


def allreduce_mean_(t: Tensor, group=None) -> Tensor:
    """In-place average over ranks (local ELBO is a batch mean, so the average over equal shards is
    the global-batch gradient).  No-op for a single process."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(dist.get_world_size(group))
    return t


def allreduce_grads_(params: Iterable[torch.nn.Parameter], group=None) -> None:
    """Flat-bucket average of .grad over ranks for arbitrary modules (encoder + head + posterior)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    allreduce_mean_(flat, group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class EmaSync:
    """Every `every` steps average the EMA shadow parameters over ranks (ranks draw different
    theta / noise, so shadows drift apart unless the gradients are identical)."""

    def __init__(self, shadow: Sequence[Tensor], every: int = 100, group=None) -> None:
        self.shadow, self.every, self.group, self.step_count = list(shadow), every, group, 0

    def step(self) -> bool:
        self.step_count += 1
        if self.step_count % self.every:
            return False
        self.sync()
        return True

    def sync(self) -> None:
        if not self.shadow:
            return
        if len(self.shadow) == 1 and self.shadow[0].dtype == torch.float32 and self.shadow[0].is_contiguous():
            allreduce_mean_(self.shadow[0], self.group)  # flat shadow (optim.FusedAdamWEma.ema): averaged in place
            return
        flat = torch.cat([s.reshape(-1).to(torch.float32) for s in self.shadow])
        allreduce_mean_(flat, self.group)
        off = 0
        for s in self.shadow:
            n = s.numel()
            s.copy_(flat[off:off + n].view_as(s))
            off += n
