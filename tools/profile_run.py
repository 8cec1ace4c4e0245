#!/usr/bin/env python
"""Small fixed workload for ncu: the bench scene (cached by bench.py), N spp in the fast mode.  Run under
`ncu --metrics gpu__time_duration.sum ...` (launch list) or `ncu --set full -k regex:k_extend ...` (one kernel)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tfg_pathtracer_b200 import renderer as R, scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=2)
ap.add_argument("--tex", type=int, default=4096)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--mode", default="fast")
ap.add_argument("--hit-mode", default="key")
ap.add_argument("--workload", default="clock")
ap.add_argument("--grid", type=int, default=2237)
ap.add_argument("--wave-spp", type=int, default=0)
a = ap.parse_args()
flat, _ = bench.get_scene(a, need_dir=False)
sc = S.load_flat(flat)
cfg = dict(R.FAST if a.mode == "fast" else R.PARITY)
if a.hit_mode == "min_t":
    cfg["hit_mode"] = R.HIT_MIN_T
if a.mode == "fast":
    cfg["wave_spp"] = a.wave_spp
r = R.Renderer(**cfg).render_setup(sc)
r.render_cuda(a.spp)
st = r.stats()
print({k: st[k] for k in ("render_ms", "rays_extension", "rays_shadow_env", "kernel_launches")})
r.close()
