"""Data-parallel plumbing for the DeformConv2d path.

The path shards only along the batch: rank r of g takes samples [r*B/g, (r+1)*B/g), weights
are replicated, and a training step ends with ONE all-reduce of a flat float32 bucket holding
every parameter gradient (DCN weight/bias, offset-conv weight/bias, the host model's other
parameters).  On CUDA the all-reduce is `dcn_allreduce_sum_f32` of the C ABI (NCCL over
NVLink, dlopen'ed); `torch.distributed` is used for rendez-vous only.  With a gloo process
group (CPU tests of this host logic) the same bucket goes through `dist.all_reduce`.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def shard_range(global_batch, rank, world):
    """[begin, end) of the samples rank `rank` owns; the shards tile the batch exactly."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(global_batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class GradBucket:
    """One contiguous float32 buffer aliasing-free copy of all gradients, in parameter order."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.numel = sum(self.sizes)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)

    def pack(self):
        off = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        return self.flat

    def unpack(self):
        off = 0
        for p, n in zip(self.params, self.sizes):
            g = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n


class DcnComm:
    """NCCL communicator behind the C ABI (dcn_comm_*); one per process = one per GPU."""

    def __init__(self, rank, world, device):
        lib = _lib.load()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            _lib.check(lib.dcn_comm_unique_id(buf), "dcn_comm_unique_id")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.to(device)
        dist.broadcast(uid, 0)          # rendez-vous over the existing process group
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.dcn_comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()),
                                         ctypes.byref(self.handle)), "dcn_comm_init")
        self.world, self.device, self.lib = world, device, lib

    def allreduce_mean_(self, flat):
        stream = torch.cuda.current_stream(self.device)
        _lib.check(self.lib.dcn_allreduce_sum_f32(self.handle, ctypes.c_void_p(flat.data_ptr()),
                                                  flat.numel(), 1.0 / self.world,
                                                  ctypes.c_void_p(stream.cuda_stream)),
                   "dcn_allreduce_sum_f32")
        return flat

    def close(self):
        if self.handle:
            self.lib.dcn_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()


def allreduce_gradients(bucket, comm=None, group=None):
    """Average the gradients of `bucket.params` over all ranks with ONE collective."""
    flat = bucket.pack()
    if comm is not None:
        comm.allreduce_mean_(flat)
    else:
        world = dist.get_world_size(group)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
    bucket.unpack()
    return flat
