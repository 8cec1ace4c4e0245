#!/usr/bin/env python
"""BVH8 build on the host cores vs on the GPU: build time, node count and the quality of the tree
measured where it matters — closest-hit throughput of the same ray batch through either tree, plus the bit-identity of
the hits (the result must not depend on the tree).  One JSON line per scene.

  python tools/bvh_build_bench.py [--grid 2237] [--rays 4000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from tfg_pathtracer_b200 import renderer as R, scenes as S  # noqa: E402


def bounce_like_rays(sc, n, seed):
    """Half camera rays, half rays leaving random surface points in random directions (what bounce rays look like)."""
    rng = np.random.RandomState(seed)
    cam = S.camera_rays(sc, n // 2, seed)
    ti = rng.randint(0, len(sc.tris), n - n // 2)
    b = rng.dirichlet((1, 1, 1), len(ti)).astype(np.float32)
    p = (sc.tris["vertices"][ti] * b[:, :, None]).sum(1)
    d = rng.randn(len(ti), 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(np.concatenate([cam, np.concatenate([p + 1e-3 * d, d], 1)]), np.float32)


def run(name, sc, nrays):
    rays = bounce_like_rays(sc, nrays, 11)
    out = {"scene": name, "tris": int(len(sc.tris)), "rays": int(len(rays))}
    hits = {}
    for label, builder in (("host", R.BVH_HOST), ("device", R.BVH_DEVICE)):
        r = R.Renderer(**dict(R.FAST, bvh_builder=builder))
        t0 = time.time()
        r.render_setup(sc)
        upload_s = time.time() - t0
        st = r.stats()
        cold_ms = st["bvh_build_ms"]
        t0 = time.time()
        r.render_setup(sc)                       # again in the same process: kernels loaded, allocator warm
        upload_s = time.time() - t0
        st = r.stats()
        d_rays = r.device_alloc(rays.nbytes)
        d_hits = r.device_alloc(len(rays) * 20)
        r.device_upload(d_rays, rays)
        r.trace_device(d_rays, len(rays), d_hits)
        ms = min(r.trace_device(d_rays, len(rays), d_hits) for _ in range(3))
        h = np.zeros(len(rays), R._capi.HIT_DT)
        r.device_download(h, d_hits)
        hits[label] = h
        out[label] = {"bvh_build_ms": st["bvh_build_ms"], "bvh_build_ms_first_call": cold_ms, "scene_upload_s": upload_s, "bvh_nodes": st["bvh_nodes"], "key_slack": st["key_slack"],
                      "closest_hit_mrays_per_s": len(rays) / ms / 1e3, "hit_fraction": float((h["tri"] >= 0).mean())}
        r.device_free(d_rays); r.device_free(d_hits); r.close()
    out["hits_identical"] = bool(hits["host"].tobytes() == hits["device"].tobytes())
    out["build_speedup"] = out["host"]["bvh_build_ms"] / out["device"]["bvh_build_ms"]
    out["trace_rate_device_tree_vs_host_tree"] = out["device"]["closest_hit_mrays_per_s"] / out["host"]["closest_hit_mrays_per_s"]
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=2237, help="n x n vertices -> 2(n-1)^2 triangles; 2237 -> 9.999 M (BASELINE config 4)")
    ap.add_argument("--rays", type=int, default=4000000)
    ap.add_argument("--skip-grid", action="store_true")
    a = ap.parse_args()
    run("clock_standin", S.clock_standin(tex_res=64), a.rays)
    if not a.skip_grid:
        run("displaced_grid_%d" % a.grid, S.displaced_grid(a.grid, env_size=(256, 128)), a.rays)


if __name__ == "__main__":
    main()
