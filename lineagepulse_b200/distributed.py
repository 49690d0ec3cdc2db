"""Gene-sharded multi-GPU plumbing (one process per GPU, torch.distributed).

The path shards by gene: fitMuDisp, evalLogLikMatrix, the impulse starts and the
LRT are independent per gene and need no collective.  The per-cell drop-out fit (fitPi) sums
over ALL genes of a cell, so each lock-step BFGS round all-reduces the N x 2q block of partial
log-likelihood sums; the EM loop all-reduces one scalar per iteration.  The CUDA library calls
back into ``make_allreduce``'s closure for those; torch.distributed (NCCL over NVLink on GPUs,
gloo in the CPU tests) does the transport.
"""
from __future__ import annotations

import numpy as np


SUM_BLOCK = 16  # LP_PI_BLOCK: shards start on summation-block boundaries => results independent of the shard count


def shard_bounds(n_genes: int, world: int, align: int = SUM_BLOCK):
    """Contiguous gene ranges of near-equal size whose starts are multiples of ``align``."""
    n_blocks = (n_genes + align - 1) // align
    base, extra = divmod(n_blocks, world)
    bounds = [0]
    for r in range(world):
        bounds.append(min(n_genes, bounds[-1] + (base + (1 if r < extra else 0)) * align))
    bounds[-1] = n_genes
    return bounds


class _CudaWordView:
    """Exposes a raw device pointer (8-byte words) to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, count: int, dtype: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8" if dtype == 1 else "<f8",
                                         "data": (ptr, False), "version": 3, "strides": None}


def init_comm(engine, group=None):
    """One process per GPU (torchrun): give ``engine`` the library's own NCCL communicator.  Rank 0 creates the
    NCCL unique id, torch.distributed only carries its 128 bytes to the other ranks; every later cross-shard sum
    is an ncclAllReduce issued by the library on its own stream (no Python in the loop)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    engine.init_comm(rank, world, box[0])
    return rank, world


def make_allreduce(group=None):
    """all-reduce(sum) callback for Engine.set_collective on CUDA buffers (the host-callback alternative to
    ``init_comm``, e.g. for a transport other than NCCL).  The collective is enqueued on the stream the library
    names, on the device the buffer lives on."""
    import torch
    import torch.distributed as dist

    def allreduce(ptr: int, count: int, dtype: int, stream: int):
        dev = torch.cuda.current_device()
        try:   # the buffer's own device, not torch's current one
            from cuda import cudart  # cuda-python
            err, attr = cudart.cudaPointerGetAttributes(ptr)
            if int(err) == 0:
                dev = int(attr.device)
        except Exception:
            pass
        with torch.cuda.device(dev):
            t = torch.as_tensor(_CudaWordView(ptr, count, dtype), device=torch.device("cuda", dev))
            assert t.data_ptr() == ptr, "the all-reduce buffer was copied instead of wrapped"
            if stream:
                ext = torch.cuda.ExternalStream(stream, device=dev)
                with torch.cuda.stream(ext):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:   # legacy default stream: order explicitly
                torch.cuda.synchronize(dev)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                torch.cuda.synchronize(dev)

    return allreduce


def allreduce_host(arr: np.ndarray, group=None) -> np.ndarray:
    """Sum a host array over the shards (gloo in tests; used for per-rank metadata)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()


def global_min(value: float, group=None) -> float:
    """min over shards, e.g. of rowMeans for the logistic_ofMu offset (initialiseDropModel.R:80)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return float(t.cpu()[0])


def gather_genes(local: np.ndarray, group=None) -> np.ndarray:
    """Concatenate per-shard gene-indexed arrays (axis 0) on every rank: the all-gather of
    G x p parameter rows / LL / convergence codes at the end of a fit."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    local = np.ascontiguousarray(local)
    if world == 1:
        return local
    objs = [None] * world
    dist.all_gather_object(objs, local, group=group)
    return np.concatenate(objs, axis=0)
