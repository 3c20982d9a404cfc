"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo process group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jittor_dcn_b200 import dp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shards_tile_the_batch_exactly():
    for B in (1, 7, 16, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, e = dp.shard_range(B, r, world)
                assert 0 <= b <= e <= B
                cover += list(range(b, e))
            assert cover == list(range(B))
            sizes = [dp.shard_range(B, r, world)[1] - dp.shard_range(B, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(8, 2, 2)


def test_bucket_roundtrip():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Linear(5, 2))
    for p in m.parameters():
        p.grad = torch.randn_like(p)
    want = [p.grad.clone() for p in m.parameters()]
    bucket = dp.GradBucket(m.parameters())
    flat = bucket.pack()
    assert flat.numel() == sum(p.numel() for p in m.parameters())
    flat.mul_(2.0)
    bucket.unpack()
    for p, w in zip(m.parameters(), want):
        assert torch.equal(p.grad, 2.0 * w)


def _worker(rank, world, port, out, batch=8):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                      # replicated weights
        model = torch.nn.Sequential(torch.nn.Conv2d(2, 3, 3, padding=1), torch.nn.Flatten(),
                                    torch.nn.Linear(3 * 6 * 6, 4))
        gen = torch.Generator().manual_seed(123)  # same global batch on every rank
        x = torch.randn(batch, 2, 6, 6, generator=gen)
        y = torch.randn(batch, 4, generator=gen)
        b, e = dp.shard_range(batch, rank, world)
        # mean over the GLOBAL batch = sum over ranks of (B_r / B) * per-shard mean; equal shards: plain average
        loss = ((model(x[b:e]) - y[b:e]) ** 2).mean()
        loss.backward()
        bucket = dp.GradBucket(model.parameters())
        weight = None if batch % world == 0 else dp.shard_weight(batch, rank, world)
        dp.allreduce_gradients(bucket, weight=weight)   # ONE collective
        if rank == 0:
            ref = torch.nn.Sequential(torch.nn.Conv2d(2, 3, 3, padding=1), torch.nn.Flatten(),
                                      torch.nn.Linear(3 * 6 * 6, 4))
            ref.load_state_dict({k: v.clone() for k, v in model.state_dict().items()})
            ((ref(x) - y) ** 2).mean().backward()
            err = max(float((p.grad - q.grad).abs().max()) for p, q in zip(model.parameters(), ref.parameters()))
            out.put(err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])   # 7: uneven shards (4 + 3), weighted by B_local / B_global
def test_two_rank_allreduce_equals_full_batch_gradients(batch):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, batch)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() < 1e-6
