#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck): device BVH build of the clock stand-in's 125 k triangles,
closest-hit batch, a parity render and a fast render with 8 samples per pixel per wave, at a tiny resolution."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import make_golden as MG  # noqa: E402
from tfg_pathtracer_b200 import renderer as R, scenes as S  # noqa: E402

small = len(sys.argv) > 1 and sys.argv[1] == "small"
sc = S.cornell_box(32, env_size=(16, 8), tilt=(3.0, 7.0, 2.0), box_gap=0.002) if small else S.clock_standin(tex_res=32, xres=64, yres=36, env_size=(64, 32))
for builder in (R.BVH_DEVICE, R.BVH_HOST):
    r = R.Renderer(bvh_builder=builder, **R.PARITY).render_setup(sc)
    h = r.trace_closest(MG.ray_batch(sc, 512, 512, 256))
    r.render_cuda(2)
    print("parity", builder, int((h["tri"] >= 0).sum()), float(r.film()[..., :3].mean()))
    r.close()
f = R.Renderer(**dict(R.FAST, wave_spp=8)).render_setup(sc)
f.render_cuda(11)
print("fast", float(f.film()[..., :3].mean()), f.stats()["kernel_launches"])
f.resolve_rgba8()
f.reduce_film(0, all_passes=True)
print("reduced", float(f.film_reduced()[..., :3].mean()))
f.close()
if not small:
    # every shading path (clearcoat, anisotropy, sheen, emission, bilinear / float / packed maps) + a point light: k_shadowLight in both flavours
    zoo = S.material_zoo(xres=48, yres=27, lights=1, env_size=(32, 16))
    for cfg in (R.PARITY, R.FAST):
        z = R.Renderer(**cfg).render_setup(zoo)
        z.render_cuda(3)
        print("zoo", float(z.film()[..., :3].mean()))
        z.close()
