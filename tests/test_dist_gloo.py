"""world_size-2 gloo test (CPU) of the N>1 host logic: sample split + the single film reduce + resolve."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tfg_pathtracer_b200 import dist as D

NPIX, TOTAL_SPP = 64, 7


def fake_sample(pixel, s):
    """Stand-in for one rendered pixel-sample (what a rank's context adds to its film sums)."""
    return np.array([np.sin(pixel * 0.37 + s), (pixel % 5) * 0.25 + s * 0.01, 1.0 / (1 + s), 0.0], np.float32)


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, stride, local = D.sample_plan(TOTAL_SPP, rank, world)
    sums = torch.zeros(NPIX * 4)
    counts = torch.zeros(NPIX, dtype=torch.int32)
    for k in range(local):
        s = off + k * stride
        for p in range(NPIX):
            if (p + s) % 11 == 0:          # a NaN-dropped sample: neither summed nor counted (S/kernel.cu:449)
                continue
            sums[4 * p:4 * p + 4] += torch.from_numpy(fake_sample(p, s))
            counts[p] += 1
    D.reduce_film(sums, counts, 0)
    if rank == 0:
        out.put((sums.numpy().copy(), counts.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_plan_partitions_all_samples():
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 5, 8, 1000):
            seen = []
            for r in range(world):
                off, stride, local = D.sample_plan(total, r, world)
                seen += [off + k * stride for k in range(local)]
            assert sorted(seen) == list(range(total))


def test_two_rank_film_reduce_matches_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums, counts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_s = np.zeros(NPIX * 4, np.float32)
    ref_c = np.zeros(NPIX, np.int32)
    for s_ in range(TOTAL_SPP):
        for p in range(NPIX):
            if (p + s_) % 11 == 0:
                continue
            ref_s[4 * p:4 * p + 4] += fake_sample(p, s_)
            ref_c[p] += 1
    assert (counts == ref_c).all()
    np.testing.assert_allclose(sums, ref_s, rtol=1e-6, atol=1e-6)
    img = D.resolve(sums, counts)
    assert img.shape == (NPIX, 4) and (img[:, 3] == 1).all()
    np.testing.assert_allclose(img[:, :3], ref_s.reshape(-1, 4)[:, :3] / ref_c[:, None], rtol=1e-5, atol=1e-6)
    t = D.resolve(torch.from_numpy(sums), torch.from_numpy(counts)).numpy()
    np.testing.assert_allclose(t, img, rtol=1e-5, atol=1e-6)
