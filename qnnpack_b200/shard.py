"""Batch-dimension data parallelism for the q8 hot path.

The reference parallelises over (image, tile) inside one process (src/operator-run.c:797-802, 675-679)
and has no cross-image dependence, so images shard freely: rank r of P owns a contiguous slab of the
NHWC batch and holds a full replica of the packed weights.  The only collective is the one-time
replication of the model parameters (uint8 kernels, int32 biases) from rank 0 (NCCL over NVLink on GPUs;
gloo in the CPU tests), after which every rank creates — plans and packs — its own operators; the
steady-state run has none.  (Raw parameters rather than packed blobs are replicated because the packing plan
depends on the values: bias magnitude selects the number of bias UMMA steps, the weight range selects the
depthwise operand split, ... — every rank must derive the plan from the same numbers.)
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


def replicate_params_from_rank0(arrays, device=None, group=None):
    """One broadcast for a whole model: `arrays` (NumPy, any dtype, identical shapes on every rank) are
    overwritten in place with rank 0's contents.  `device` = torch device the collective runs on (a CUDA
    device for NCCL, None/cpu for gloo).  Returns the bytes moved."""
    import numpy as np
    import torch
    import torch.distributed as dist

    sizes = [a.nbytes for a in arrays]
    flat = np.empty(sum(sizes), dtype=np.uint8)
    off = 0
    for a, n in zip(arrays, sizes):
        flat[off:off + n] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        off += n
    t = torch.from_numpy(flat)
    if device is not None and str(device) != "cpu":
        t = t.to(device)
    dist.broadcast(t, src=0, group=group)
    flat = t.cpu().numpy()
    off = 0
    for a, n in zip(arrays, sizes):
        a[...] = flat[off:off + n].view(a.dtype).reshape(a.shape)
        off += n
    return int(sum(sizes))


class DeviceBytes:
    """A raw device allocation (pointer, size) exposed through __cuda_array_interface__ so that
    torch.as_tensor() can wrap library-owned memory (e.g. qnnp_cuda_operator_packed_weights) for NCCL."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
