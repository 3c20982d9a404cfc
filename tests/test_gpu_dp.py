"""Data-parallel parity on real GPUs (SURVEY.md 8e "Verification"): the DCN layer's parameter gradients after the
all-reduce at g = 2 GPUs equal the single-GPU full-batch gradients to fp32 re-association tolerance — through both
exchanges of the C ABI: `dcn_allreduce_sum_f32` (NCCL) and `dcn_p2p_allreduce_sum_f32` (one kernel over NVLink peer
memory), the latter also replayed from a CUDA graph.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

GLOBAL_BATCH = 10      # shards 5 + 5 at g = 2; 7 -> 4 + 3 (uneven, weighted by B_local / B_global)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _layer_and_data(dev, batch):
    import jittor_dcn_b200 as dcn
    torch.manual_seed(5)                       # replicated weights, same global batch on every rank
    layer = dcn.TorchDeformConv2d(64, 64, 3, 1, 1)
    with torch.no_grad():
        layer.offset_conv.weight.normal_(0, 0.02)
        layer.offset_conv.bias.normal_(0, 1.0)
        layer.bias.normal_(0, 0.1)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(batch, 64, 16, 16, generator=g)
    gout = torch.randn(batch, 64, 16, 16, generator=g)
    return layer.to(dev), x.to(dev), gout.to(dev)


def _worker(rank, world, port, batch, out):
    import torch.distributed as dist
    from jittor_dcn_b200 import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        layer, x, gout = _layer_and_data(dev, batch)
        b, e = dp.shard_range(batch, rank, world)
        loss = (layer(x[b:e]) * gout[b:e]).sum() / (e - b)          # mean over the rank's shard
        loss.backward()
        weight = dp.shard_weight(batch, rank, world)
        local = [p.grad.clone() for p in layer.parameters()]
        results = {}
        bucket = dp.GradBucket(layer.parameters())
        for kind in ("nccl", "p2p", "p2p_graph"):
            for p, g in zip(layer.parameters(), local):
                p.grad.copy_(g)
            if kind == "nccl":
                comm = dp.DcnComm(rank, world, dev)
                dp.allreduce_gradients(bucket, comm, weight=weight)
            else:
                comm = dp.DcnP2P(rank, world, dev, bucket.numel)
                if kind == "p2p":
                    dp.allreduce_gradients(bucket, comm, weight=weight)
                else:
                    # the exchange recorded in a CUDA graph and replayed twice (fresh inputs each time)
                    flat = bucket.pack()
                    keep = flat.clone()
                    side = torch.cuda.Stream(dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        comm.allreduce_sum_(flat, 1.0)               # eager warm-up on the capture stream
                    torch.cuda.current_stream(dev).wait_stream(side)
                    torch.cuda.synchronize(dev)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        comm.allreduce_sum_(flat, 1.0)
                    for _ in range(2):
                        flat.copy_(keep).mul_(weight)
                        graph.replay()
                    torch.cuda.synchronize(dev)
                    bucket.unpack()
            torch.cuda.synchronize(dev)
            results[kind] = [p.grad.clone() for p in layer.parameters()]
            dist.barrier()
            comm.close()
        if rank == 0:
            ref, xr, gr = _layer_and_data(dev, batch)
            ((ref(xr) * gr).sum() / batch).backward()
            errs = {}
            for kind, grads in results.items():
                errs[kind] = max(float((g - p.grad).abs().max() / p.grad.abs().max())
                                 for g, p in zip(grads, ref.parameters()))
            # all ranks must hold bit-identical reduced gradients: checked against rank 1 below
            out.put(errs)
        flat_all = torch.cat([g.reshape(-1) for g in results["p2p"]])
        gathered = [torch.empty_like(flat_all) for _ in range(world)]
        dist.all_gather(gathered, flat_all)
        if rank == 0:
            out.put(bool(all(torch.equal(gathered[0], t) for t in gathered[1:])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [GLOBAL_BATCH, 7])
def test_two_gpu_allreduce_equals_full_batch_gradients(batch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    errs = out.get()
    for kind, err in errs.items():
        assert err < 1e-4, (kind, err)
    assert out.get() is True, "ranks disagree bitwise after the peer-memory all-reduce"
