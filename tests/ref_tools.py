"""Runs THE REFERENCE RENDERER's own CUDA build (oracle/_ref/eleven_ref_headless_{precise,fast}: its loader, its BVH, its
kernel.cu, compiled from /root/reference by oracle/Makefile) on the GPU box, from a scene directory in its own on-disk format.
TEST INFRASTRUCTURE (the checker), used by tests/ and the fixture generators under tests/golden/ only.  The binaries are
git-ignored build products that travel with the snapshot; where they are absent the callers skip."""
import json
import os
import shutil
import subprocess
import tempfile

import numpy as np

from tfg_pathtracer_b200 import scenes as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_binary(flavour="precise"):
    return os.path.join(ROOT, "oracle", "_ref", "eleven_ref_headless_" + flavour)


def have_ref(flavour="precise"):
    return os.path.exists(ref_binary(flavour))


class RefRun:
    """Result of one reference run: film passes (H, W, 4), path counts, beauty snapshots {n: (H, W, 4)}, the timing record, and the
    scene EXACTLY as the reference's loader produced it (MikkTSpace tangents, stb texel decode) as a SceneData for our side."""

    def __init__(self, info, passes, pathcount, snapshots, scene):
        self.info, self.passes, self.pathcount, self.snapshots, self.scene = info, passes, pathcount, snapshots, scene


def run_reference(scene: S.SceneData, spp, flavour="precise", env_color=None, snapshots=(), external_textures=False, workdir=None, keep=False,
                  timeout=3000):
    work = workdir or tempfile.mkdtemp(prefix="eleven_ref_")
    try:
        S.write_reference_scene_dir(scene, work, env_color=env_color)
        prefix, dump = os.path.join(work, "ref"), os.path.join(work, "scene.flat")
        cmd = [ref_binary(flavour), work, str(int(spp)), prefix, "--dump-scene", dump]
        if external_textures:
            cmd.append("--external-textures")
        snaps = [int(n) for n in snapshots if 0 < int(n) < int(spp)]
        if snaps:
            cmd += ["--snapshots", ",".join(map(str, snaps))]
        p = subprocess.run(cmd, capture_output=True, text=True, cwd=work, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("reference run failed rc=%d: %s" % (p.returncode, (p.stderr or p.stdout)[-800:]))
        info = json.loads(open(prefix + ".json").read())
        W, H = info["width"], info["height"]
        passes = {k: np.fromfile(prefix + "." + k + ".f32", np.float32).reshape(H, W, 4) for k in ("beauty", "normal", "tangent", "bitangent")}
        pc = np.fromfile(prefix + ".pathcount.i32", np.int32)
        shots = {n: np.fromfile("%s.beauty@%d.f32" % (prefix, n), np.float32).reshape(H, W, 4) for n in snaps}
        return RefRun(info, passes, pc, shots, S.load_flat(dump))
    finally:
        if not keep and workdir is None:
            shutil.rmtree(work, ignore_errors=True)


def within(a, b, atol=1e-3, rtol=1e-3):
    return float((np.abs(a - b) <= atol + rtol * np.abs(b)).all(-1).mean())


def rmse(a, b):
    return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))
