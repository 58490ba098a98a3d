"""Host-side multi-rank plumbing (one process per GPU, torch.distributed for rendezvous only).

The path has ONE exchange per step: a sum-all-reduce of the partial gradient sums.
On GPUs it runs as our own one-shot NVLink kernel (csrc/comm.cu) once the ranks have swapped CUDA IPC
handles through `torch.distributed.all_gather_object`; NCCL through torch is the fallback callback.
Everything here is pure host logic and is exercised under gloo on CPU by tests/test_parallel_gloo.py.
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous near-equal split of range(total): returns (start, length)."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def sample_shard(M: int, rank: int, world: int):
    """M-axis sharding: rank r evaluates Monte-Carlo samples [m0, m0 + M_local) of the global M;
    eps is addressed by the GLOBAL sample index, so results do not depend on the world size."""
    return shard_range(M, rank, world)


def row_shard(n: int, rank: int, world: int, align: int = 1):
    """n-axis sharding of the data rows; starts are multiples of `align`."""
    blocks = -(-n // align)
    b0, nb = shard_range(blocks, rank, world)
    r0 = min(b0 * align, n)
    return r0, min((b0 + nb) * align, n) - r0


def broadcast_key(key: int, src: int = 0) -> int:
    """All ranks must draw eps from the same Philox key."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([key & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src)
    return int(t.item())


def rank_batches(perm_batches, rank: int, world: int):
    """Weak-scaling minibatches (config 5): of the epoch's batch list, rank r takes batches r, r+world, ...;
    one global step consumes `world` consecutive batches (global batch = world * batchsize)."""
    usable = len(perm_batches) - len(perm_batches) % world
    return [perm_batches[i] for i in range(rank, usable, world)]


def allgather_bytes(b: bytes):
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, b)
    return out


def make_torch_allreduce():
    """Fallback exchange: NCCL all-reduce through torch.distributed on the library's stream."""
    import torch
    import torch.distributed as dist

    def fn(dev_ptr: int, count: int, stream: int):
        class _Buf:   # __cuda_array_interface__ view of the library-owned buffer
            __cuda_array_interface__ = dict(shape=(count,), typestr="<f4", data=(dev_ptr, False), version=2)
        ext = torch.cuda.ExternalStream(stream)
        with torch.cuda.stream(ext):
            t = torch.as_tensor(_Buf(), device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return fn


def connect(ctx, max_floats: int, native: bool = True):
    """Wire `ctx` into the process group this rank belongs to."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    if native:
        ctx.connect_peers(max_floats, rank, world, allgather_bytes)
    else:
        ctx.set_allreduce(make_torch_allreduce(), rank, world)
