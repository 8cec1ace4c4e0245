"""One-process-per-GPU driver logic: the scene is replicated, samples are split across ranks, and the
per-GPU film SUMS (+ per-pixel sample counts) are summed with ONE reduce to rank 0 (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  torch.distributed is plumbing only; there is no data-path collective besides that reduce.

The reference has no multi-GPU path (cudaSetDevice(0), S/kernel.cu:604); its running-mean film (S/kernel.cu:451-477)
cannot be combined across devices, which is why the film is kept as sums here.
"""
from __future__ import annotations

import numpy as np


def sample_plan(total_spp: int, rank: int, world: int):
    """Global sample s is rendered by rank s % world.  Returns (sample_offset, sample_stride, local_spp)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    local = (total_spp - rank + world - 1) // world if total_spp > rank else 0
    return rank, world, local


class DeviceArray:
    """Zero-copy __cuda_array_interface__ view of a buffer owned by the C-ABI context (eleven_film_*_device)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def film_tensors(renderer, device):
    """torch views (no copy) of the context's BEAUTY sums (float32, W*H*4) and sample counts (int32, W*H)."""
    import torch
    from .renderer import PASS_BEAUTY
    sums = torch.as_tensor(DeviceArray(*renderer.film_sums_ptr(PASS_BEAUTY), "<f4"), device=device)
    counts = torch.as_tensor(DeviceArray(*renderer.film_counts_ptr(), "<i4"), device=device)
    return sums, counts


def reduce_film(sums, counts, dst: int = 0, group=None):
    """The one exchange step: sum film sums and counts onto rank `dst` (in place there)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(sums, dst, group=group)
        dist.reduce(counts, dst, group=group)
    return sums, counts


def resolve(sums, counts):
    """mean = sum / count per pixel, alpha = 1 (the getBuffers contract, S/kernel.cu:137,461-463).  numpy or torch."""
    s = sums.reshape(-1, 4)
    c = counts.reshape(-1, 1)
    if isinstance(s, np.ndarray):
        out = np.where(c > 0, s / np.maximum(c, 1), 0).astype(np.float32)
        out[:, 3] = 1.0
        return out
    import torch
    out = torch.where(c > 0, s / c.clamp(min=1).to(s.dtype), torch.zeros_like(s))
    out[:, 3] = 1.0
    return out
