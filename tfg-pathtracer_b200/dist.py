"""One-process-per-GPU driver logic: the scene is replicated, samples are split across ranks, and the per-GPU
film records (per-pixel SUMS whose .w carries the accepted-sample count) are summed with ONE reduce to the root.

On the GPU box that reduce is native: `eleven_reduce_film` issues one `ncclReduce` over NVLink on the context's render
stream, into a buffer separate from the local film (Renderer.reduce_film / film_reduced).  torch.distributed is plumbing
only — it carries the 128-byte NCCL unique id to the ranks (`init_comm`) and provides the barriers of bench.py.
`reduce_records` is the same exchange step on host tensors; it is what the world_size-2 gloo tests on CPU exercise and
what a caller that brings its own collective would do with `eleven_film_sums_device`.

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


def job_plan(total_spp: int, rank: int, world: int, wave_spp: int = 16):
    """Strong-scaling split of a FIXED job (`eleven <scene> 1000 out.bmp --gpus N`): the job's samples are cut into waves of
    `wave_spp` (what a context renders per wave at full speed) and rank r gets a contiguous run of whole waves, so that no rank
    renders ragged 8 + 4 + 1-sample waves: with s % world, 1000 spp on 8 GPUs is 125 per GPU = 7 full waves + 3 small ones; here it
    is 8 waves on seven GPUs and 7 (the last one half full) on the eighth.  The counter RNG is keyed by the GLOBAL sample index, so
    the union over ranks — the image — is the same for any split.  Returns (sample_offset, sample_stride = 1, local_spp)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    waves = (total_spp + wave_spp - 1) // wave_spp
    w0 = rank * (waves // world) + min(rank, waves % world)
    w1 = w0 + waves // world + (1 if rank < waves % world else 0)
    s0, s1 = min(total_spp, w0 * wave_spp), min(total_spp, w1 * wave_spp)
    return s0, 1, s1 - s0


def init_comm(renderer, rank: int, world: int, src: int = 0, group=None):
    """Creates the film-reduce communicator of `renderer` (eleven_comm_init_rank): rank `src` draws the NCCL unique id,
    torch.distributed (any backend) hands it to the others.  Collective."""
    import torch.distributed as dist
    box = [renderer.comm_unique_id() if rank == src else None]
    if world > 1:
        dist.broadcast_object_list(box, src=src, group=group)
    renderer.comm_init_rank(box[0], world, rank)


def reduce_records(records, dst: int = 0, group=None):
    """The exchange step on a host/torch tensor of film records (n x 4 float32: sum.xyz, count): returns the sum over ranks
    on rank `dst` as a NEW tensor (the local records are left untouched, like eleven_reduce_film: the step can be repeated
    while rendering goes on without double-counting earlier samples); other ranks get None."""
    import torch.distributed as dist
    out = records.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(out, dst, group=group)
        if dist.get_rank(group) != dst:
            return None
    return out


def resolve(records):
    """mean = sum.xyz / count per pixel, alpha = 1 (the getBuffers contract, S/kernel.cu:137,461-463).  numpy or torch."""
    s = records.reshape(-1, 4)
    c = s[:, 3:4]
    if isinstance(s, np.ndarray):
        out = np.where(c > 0, s / np.maximum(c, 1), 0).astype(np.float32)
        out[:, 3] = 1.0
        return out
    import torch
    out = torch.where(c > 0, s / c.clamp(min=1), torch.zeros_like(s))
    out[:, 3] = 1.0
    return out
