#!/usr/bin/env python
"""Fixture generator (runs on the GPU box; needs oracle/_ref): converged REFERENCE renders for the convergence tests.

  convergence_clock.npz   ClockCC0 stand-in at 480x270 (256^2 maps, 1024x512 HDRI), THE REFERENCE'S CUDA BUILD (precise flavour) at
                          16 384 spp = the "converged reference" of BASELINE.json's north_star part 2, with its own film snapshots at
                          16 / 64 / 256 / 1000 spp (one run: the reference's N-spp image is a prefix of its converged one) -> RMSE_ref(N),
                          plus the 16-spp image itself so that the test can check on ITS box that the reference still renders these bits
  fullframe_1000spp.npz   (--fullframe) the 1920x1080 benchmark scene at 1000 spp by the reference (shipping -use_fast_math flavour, the
                          one the bench times), stored as 4x4-block means (480x270x3) for the RMSE-between-renderers test

usage (repo root, GPU box):  python tests/golden/make_convergence.py [--spp 16384] [--fullframe]
outputs land in gpurun_out/ (copy them to tests/golden/ to commit)."""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)
import ref_tools as RT  # noqa: E402
from tfg_pathtracer_b200 import scenes as S  # noqa: E402

LEVELS = (16, 64, 256, 1000)


def convergence_scene():
    return S.clock_standin(tex_res=256, xres=480, yres=270, env_size=(1024, 512))


def block_mean(img, k=4):
    H, W = img.shape[:2]
    return img[:H // k * k, :W // k * k].reshape(H // k, k, W // k, k, -1).mean((1, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=16384)
    ap.add_argument("--fullframe", action="store_true")
    ap.add_argument("--fullframe-spp", type=int, default=1000)
    ap.add_argument("--skip-convergence", action="store_true")
    a = ap.parse_args()
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    if not a.skip_convergence:
        t0 = time.time()
        ref = RT.run_reference(convergence_scene(), a.spp, "precise", snapshots=LEVELS, timeout=6000)
        conv = ref.passes["beauty"][..., :3]
        rm = {n: RT.rmse(ref.snapshots[n][..., :3], conv) for n in LEVELS}
        np.savez_compressed(os.path.join(out, "convergence_clock.npz"), converged=conv, converged_spp=a.spp, levels=np.array(LEVELS),
                            rmse_ref=np.array([rm[n] for n in LEVELS]), mean_ref=np.array([ref.snapshots[n][..., :3].mean() for n in LEVELS]),
                            ref16=ref.snapshots[16][..., :3], flavour="precise", info=json.dumps(ref.info))
        print(json.dumps({"convergence": {"spp": a.spp, "wall_s": time.time() - t0, "rmse_ref": rm, "mean": float(conv.mean()), "info": ref.info}}))
    if a.fullframe:
        import argparse as ap2
        import bench
        t0 = time.time()
        flat, _ = bench.get_scene(ap2.Namespace(tex=4096, width=1920, height=1080, workload="clock", grid=0), need_dir=False)
        sc = S.load_flat(flat)
        sc.object_names, sc.material_names = ["clock", "table", "plant"], ["clock_mat", "table_mat", "plant_mat"]
        ref = RT.run_reference(sc, a.fullframe_spp, "fast", external_textures=True, timeout=6000)
        img = ref.passes["beauty"][..., :3]
        np.savez_compressed(os.path.join(out, "fullframe_1000spp.npz"), block4=block_mean(img).astype(np.float32), spp=a.fullframe_spp, mean=float(img.mean()),
                            flavour="fast", info=json.dumps(ref.info))
        print(json.dumps({"fullframe": {"spp": a.fullframe_spp, "wall_s": time.time() - t0, "mean": float(img.mean()), "info": ref.info}}))


if __name__ == "__main__":
    main()
