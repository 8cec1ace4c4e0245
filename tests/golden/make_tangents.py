#!/usr/bin/env python
"""Golden vectors for the product loader's tangent space (tfg-pathtracer_b200/host/tangent_space.cpp).

Runs THE REFERENCE'S OWN LOADER — S/ObjLoader.hpp + its vendored mikktspace, compiled from /root/reference into
oracle/_ref/eleven_ref_headless_precise (oracle/Makefile), dump-only mode (no GPU needed) — on scene directories written by our
generators, and stores the tangents + handedness it produced:  tests/golden/tangents.npz.  The scenes: the Cornell box, the displaced
grid, the ClockCC0 stand-in (a seeded subsample of 6 000 triangles is stored, all 125 281 are compared when this script runs) and
`tangent_torture` (scenes.py: UV seams, mirrored UV islands, degenerate triangles, zero-area UV mappings, a butterfly edge).
Needs /root/reference at build time only; the committed fixture travels.  Usage: python tests/golden/make_tangents.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
import make_golden as MG                       # noqa: E402
from tfg_pathtracer_b200 import scenes as S    # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "eleven_ref_headless_precise")
CLI = os.path.join(ROOT, "tfg-pathtracer_b200", "host", "eleven")


def tangent_scenes():
    g = MG.golden_scenes()
    return {"cornell": g["cornell"], "grid": g["grid"], "clock": g["clock"], "torture": S.tangent_torture()}


def load_with(binary_args, scene, work):
    S.write_reference_scene_dir(scene, work)
    out = os.path.join(work, "dump.flat")
    p = subprocess.run([a.replace("@DIR", work).replace("@OUT", out) for a in binary_args], capture_output=True, text=True, cwd=work)
    if p.returncode != 0:
        raise RuntimeError((p.stderr or p.stdout)[-600:])
    return S.load_flat(out).tris


def reference_tris(scene):
    with tempfile.TemporaryDirectory(prefix="eleven_tan_") as w:
        return load_with([REF, "@DIR", "0", os.path.join(w, "o"), "--dump-scene", "@OUT"], scene, w)


def product_tris(scene):
    with tempfile.TemporaryDirectory(prefix="eleven_tan_") as w:
        return load_with([CLI, "--dump-flat", "@DIR", "@OUT"], scene, w)


def subsample(name, n):
    return np.arange(n) if n <= 8000 else np.sort(np.random.RandomState(len(name) + 17).choice(n, 6000, replace=False))


if __name__ == "__main__":
    out = {}
    for name, sc in tangent_scenes().items():
        ref, ours = reference_tris(sc), product_tris(sc)
        for f in ("vertices", "uv", "normals"):
            assert (ref[f].view(np.uint32) == ours[f].view(np.uint32)).all(), (name, f)
        d = np.abs(ref["tangents"] - ours["tangents"]).max(-1)
        print("%-8s %7d triangles: tangents max |d| %.3e, corners bit-equal %.4f, handedness equal %.6f, negative handedness %.3f, default frames %d"
              % (name, len(ref), d.max(), (ref["tangents"].view(np.uint32) == ours["tangents"].view(np.uint32)).all(-1).mean(),
                 (ref["tangentsSign"] == ours["tangentsSign"]).mean(), (ref["tangentsSign"] < 0).mean(),
                 int(((ref["tangents"] == np.array([1, 0, 0], np.float32)).all(-1)).sum())))
        idx = subsample(name, len(ref))
        out[name + "_index"] = idx.astype(np.int32)
        out[name + "_vertices"] = ref["vertices"][idx]
        out[name + "_tangents"] = ref["tangents"][idx]
        out[name + "_sign"] = ref["tangentsSign"][idx]
        out[name + "_count"] = np.int64(len(ref))
    np.savez_compressed(os.path.join(HERE, "tangents.npz"), **out)
    print("wrote tests/golden/tangents.npz (%.0f KB)" % (os.path.getsize(os.path.join(HERE, "tangents.npz")) / 1024))
