"""world_size-2 gloo tests (CPU) of the N>1 host logic: sample split, the single reduce of film records (sums + count in .w)
into a buffer separate from the local film, repeated across steps, and the resolve."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tfg_pathtracer_b200 import dist as D

NPIX, SPP_PER_STEP, STEPS = 64, 7, 2


def fake_sample(pixel, s):
    """Stand-in for one rendered pixel-sample (what a rank's context adds to its film record): rgb, and 1 sample counted."""
    return np.array([np.sin(pixel * 0.37 + s), (pixel % 5) * 0.25 + s * 0.01, 1.0 / (1 + s), 1.0], np.float32)


def dropped(pixel, s):
    return (pixel + s) % 11 == 0          # a NaN-dropped sample: neither summed nor counted (S/kernel.cu:449)


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the id hand-off of init_comm, without a GPU: a stand-in "renderer" records what it was given
    class FakeRenderer:
        def comm_unique_id(self):
            return bytes(range(128))
        def comm_init_rank(self, uid, n, r):
            self.got = (uid, n, r)
    fr = FakeRenderer()
    D.init_comm(fr, rank, world)
    assert fr.got == (bytes(range(128)), world, rank)
    records = torch.zeros(NPIX, 4)
    per_step = []
    for step in range(STEPS):                                  # progressive: the film keeps accumulating, every step reduces it
        off, stride, local = D.sample_plan(SPP_PER_STEP, rank, world)
        for k in range(local):
            s = step * SPP_PER_STEP + off + k * stride
            for p in range(NPIX):
                if not dropped(p, s):
                    records[p] += torch.from_numpy(fake_sample(p, s))
        before = records.clone()
        red = D.reduce_records(records, 0)
        assert torch.equal(records, before), "the local film must not be touched by the reduce"
        assert (red is None) == (rank != 0)
        if rank == 0:
            per_step.append(red.numpy().copy())
    if rank == 0:
        out.put(per_step)
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


def test_job_plan_partitions_all_samples_in_whole_waves():
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 5, 16, 100, 1000, 4096):
            seen, most = [], 0
            for r in range(world):
                off, stride, local = D.job_plan(total, r, world, 16)
                assert stride == 1 and (off % 16 == 0 or local == 0)
                seen += list(range(off, off + local)); most = max(most, local)
            assert sorted(seen) == list(range(total))
            assert most <= 16 * -(-(-(-total // 16)) // world)            # nobody renders more than ceil(waves / world) waves
    assert [D.job_plan(1000, r, 8)[2] for r in range(8)] == [128] * 7 + [104]


def test_two_rank_film_reduce_matches_single_rank_at_every_step():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    per_step = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.zeros((NPIX, 4), np.float32)
    for step in range(STEPS):
        for s_ in range(step * SPP_PER_STEP, (step + 1) * SPP_PER_STEP):
            for p in range(NPIX):
                if not dropped(p, s_):
                    ref[p] += fake_sample(p, s_)
        got = per_step[step]
        assert (got[:, 3] == ref[:, 3]).all(), "sample counts after step %d (no double counting of earlier steps)" % step
        np.testing.assert_allclose(got[:, :3], ref[:, :3], rtol=1e-6, atol=1e-6)
    img = D.resolve(per_step[-1])
    assert img.shape == (NPIX, 4) and (img[:, 3] == 1).all()
    np.testing.assert_allclose(img[:, :3], ref[:, :3] / ref[:, 3:4], rtol=1e-5, atol=1e-6)
    t = D.resolve(torch.from_numpy(per_step[-1])).numpy()
    np.testing.assert_allclose(t, img, rtol=1e-5, atol=1e-6)
    z = D.resolve(np.zeros((3, 4), np.float32))
    assert (z[:, :3] == 0).all() and (z[:, 3] == 1).all()
