"""Batch-dimension data parallelism for the q8 hot path.

The reference parallelises over (image, tile) inside one process (src/operator-run.c:797-802, 675-679)
and has no cross-image dependence, so images shard freely: rank r of P owns a contiguous slab of the
NHWC batch and holds a full replica of the packed weights.  The only collective is the one-time
replication of the packed weight/bias blobs from rank 0 (NCCL over NVLink on GPUs; gloo in the CPU
tests); the steady-state run has none.
"""
from __future__ import annotations


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """[begin, end) of the images owned by `rank`; the first total % world ranks get one extra."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def replicate_from_rank0(tensors, group=None):
    """Broadcast every tensor (packed weights, folded biases) from rank 0 in place. Returns bytes moved."""
    import torch.distributed as dist

    moved = 0
    for t in tensors:
        dist.broadcast(t, src=0, group=group)
        moved += t.numel() * t.element_size()
    return moved


class DeviceBytes:
    """A raw device allocation (pointer, size) exposed through __cuda_array_interface__ so that
    torch.as_tensor() can wrap library-owned memory (e.g. qnnp_cuda_operator_packed_weights) for NCCL."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
