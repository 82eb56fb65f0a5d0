"""World-size-2 gloo tests (CPU) of the data-parallel plumbing: shard ranges, flat-bucket gradient
average, generic gradient average and EMA sync."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from viforsdes_b200.dist import EmaSync, FlatBucket, allreduce_grads_, allreduce_mean_, shard_range


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # flat bucket: views alias one buffer; one all-reduce averages every gradient
        b = FlatBucket([(3, 5), (7,), (2, 2)], "cpu")
        for i, v in enumerate(b.views):
            v.fill_(float(rank + 1) * (i + 1))
        b.allreduce_mean_()
        for i, v in enumerate(b.views):
            assert torch.allclose(v, torch.full_like(v, 1.5 * (i + 1)))
        assert all(v.data_ptr() >= b.flat.data_ptr() for v in b.views)
        # extra tail slot: the ELBO scalar rides in the same collective as the gradients
        b2 = FlatBucket([(4, 3), (5,)], "cpu", extra=1)
        for v in b2.views:
            v.fill_(float(rank))
        torch.sum(torch.full((6,), float(rank + 1)), dim=0, keepdim=True, out=b2.extra)
        b2.allreduce_mean_()
        assert b2.extra.shape == (1,) and b2.extra.item() == 9.0  # mean of 6 and 12
        assert all(torch.allclose(v, torch.full_like(v, 0.5)) for v in b2.views)
        # sum of shard gradients == single-process gradient on the concatenated batch (local loss is a batch mean)
        torch.manual_seed(0)
        lin = torch.nn.Linear(4, 3)
        x = torch.arange(32, dtype=torch.float32).reshape(8, 4) / 10
        lo, hi = shard_range(8, rank, world)
        lin(x[lo:hi]).pow(2).mean().backward()
        allreduce_grads_(lin.parameters())
        ref = torch.nn.Linear(4, 3)
        ref.load_state_dict(lin.state_dict())
        ref(x).pow(2).mean().backward()
        assert torch.allclose(lin.weight.grad, ref.weight.grad, atol=1e-6)
        assert torch.allclose(lin.bias.grad, ref.bias.grad, atol=1e-6)
        # EMA sync every 2 steps
        shadow = [torch.full((4,), float(rank)), torch.full((2, 2), 10.0 * rank)]
        ema = EmaSync(shadow, every=2)
        assert ema.step() is False and shadow[0][0].item() == float(rank)
        assert ema.step() is True
        assert torch.allclose(shadow[0], torch.full((4,), 0.5)) and torch.allclose(shadow[1], torch.full((2, 2), 5.0))
        t = torch.tensor([float(rank)])
        assert allreduce_mean_(t).item() == 0.5
        # FlatParameters: the gradient views of a module ARE the bucket, one collective averages them in place
        from viforsdes_b200.optim import FlatParameters

        torch.manual_seed(1)
        net = torch.nn.Linear(4, 3)
        flat = FlatParameters([list(net.parameters())])
        net(x[lo:hi]).pow(2).mean().backward()
        allreduce_mean_(flat.grads)
        ref2 = torch.nn.Linear(4, 3)
        with torch.no_grad():
            ref2.weight.copy_(net.weight)
            ref2.bias.copy_(net.bias)
        ref2(x).pow(2).mean().backward()
        assert torch.allclose(net.weight.grad, ref2.weight.grad, atol=1e-6)
        assert torch.allclose(net.bias.grad, ref2.bias.grad, atol=1e-6)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_batch():
    for total in (1, 7, 8, 8192):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world, allow_uneven=True) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            # equal-weight averaging of per-rank means is the global mean only with the shard weights applied
            from viforsdes_b200.dist import shard_weight
            assert abs(sum(shard_weight(total, r, world) for r in range(world)) - world) < 1e-12
            if total % world:
                with pytest.raises(ValueError):
                    shard_range(total, 0, world)
            else:
                assert spans == [shard_range(total, r, world) for r in range(world)]


@pytest.mark.timeout(120)
def test_two_rank_gloo():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
