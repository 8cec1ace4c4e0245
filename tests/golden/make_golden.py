#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ by running THE REFERENCE'S OWN CODE.

Runs only in the build container (needs /root/reference and oracle/_ref/ref_host_vectors, built by
`make -C oracle`).  The fixtures pin oracle/eleven_oracle.cpp (tests/test_oracle_golden.py):

  xorwow.npz    curand_init(0, subseq, 0) states + 32 curand_uniform draws for 9 subsequences (toolkit header, host path)
  texel.npz     the 2 x 256 LDR decode tables from stb_image's patched loader (sRGB gamma 2.2f, linear 1.0f)
  disney.npz    DisneyEval / DisneyPdf / DisneySample on 1500 seeded (material, V, L, r) records
  scene_<name>.npz  per scene: sha256 of the reference BVH (boxes, from/to/depth, triIndices), HDRI CDF digest +
                radianceSum + NEE samples (texel, direction, pdf), environment lookups, and the reference's own
                BVH::transverse Hit for a seeded ray batch (position, normal, tangent, bitangent, tu, tv, objectID)
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tfg_pathtracer_b200 import scenes as S  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_host_vectors")


def read_container(path):
    out = {}
    raw = open(path, "rb").read()
    pos = 0
    while pos < len(raw):
        nl = int(np.frombuffer(raw, "<u4", 1, pos)[0]); pos += 4
        name = raw[pos:pos + nl].decode(); pos += nl
        dt = int(np.frombuffer(raw, "<u4", 1, pos)[0]); pos += 4
        cnt = int(np.frombuffer(raw, "<u8", 1, pos)[0]); pos += 8
        out[name] = np.frombuffer(raw, ["<f4", "<i4", "<u4"][dt], cnt, pos).copy(); pos += 4 * cnt
    return out


def run_ref(*args):
    subprocess.check_call([REF_BIN] + list(args), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_scenes():
    """The parity scenes: small enough for the CPU suite, covering big flat triangles (cornell), smooth small
    triangles + textures + defocus (clock stand-in), a regular height field (grid) and tiling/offset env."""
    # tilted + boxes lifted 2 mm off the floor: no NaN slab misses, no coplanar equal-key ties (see scenes.cornell_box)
    c = S.cornell_box(96, env_size=(66, 33), tilt=(3.0, 7.0, 2.0), box_gap=0.002)
    k = S.clock_standin(tex_res=32, xres=160, yres=90, env_size=(128, 64))
    k.hdri.xOffset = 0.25
    g = S.displaced_grid(64, xres=96, yres=54, env_size=(64, 32))
    return {"cornell": c, "clock": k, "grid": g}


def ray_batch(scene, n_cam=1536, n_rand=1024, n_surf=512, seed=7):
    rng = np.random.RandomState(seed)
    cam = S.camera_rays(scene, n_cam, seed)
    V = scene.tris["vertices"].reshape(-1, 3)
    lo, hi = V.min(0), V.max(0)
    o = lo + (hi - lo) * (rng.rand(n_rand, 3) * 1.4 - 0.2)
    d = rng.randn(n_rand, 3)
    rnd = np.concatenate([o, d], 1)
    # rays leaving surface points (like bounce / shadow rays), including exactly-on-vertex and on-edge targets
    ti = rng.randint(0, len(scene.tris), n_surf)
    b = rng.dirichlet((1, 1, 1), n_surf)
    b[: n_surf // 8] = np.eye(3)[rng.randint(0, 3, n_surf // 8)]            # aim at vertices (ties)
    b[n_surf // 8: n_surf // 4, 2] = 0                                     # aim at edges (ties)
    b /= b.sum(1, keepdims=True)
    tgt = (scene.tris["vertices"][ti] * b[:, :, None]).sum(1)
    org = lo + (hi - lo) * rng.rand(n_surf, 3)
    surf = np.concatenate([org, tgt - org], 1)
    return np.concatenate([cam, rnd, surf]).astype(np.float32)


def disney_records(n=1500, seed=3):
    rng = np.random.RandomState(seed)
    r = np.zeros((n, 30), np.float32)
    hd = rng.rand(n, 12).astype(np.float32)
    hd[:, 5] = 1.45                     # eta (unused)
    hd[:, 6] = 0                        # transmission
    hd[: n // 5, 0] = 0                 # metallic 0
    hd[n // 5: 2 * n // 5, 1] = 1       # roughness 1 (the file-format default)
    hd[2 * n // 5: 3 * n // 5, 2:5] = 0  # no clearcoat / anisotropy
    hd[-20:, 1] = 0                     # roughness 0 -> alpha clamp
    r[:, 0:12] = hd
    r[:, 12:15] = rng.rand(n, 3) * (rng.rand(n, 1) < 0.3)          # emission
    r[:, 15:18] = rng.rand(n, 3)                                   # albedo
    r[-10:, 15:18] = 0                                             # black albedo (Cdlum == 0 branch)
    nrm = rng.randn(n, 3); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[:30] = (0, 1, 0)                                           # degenerate createBasis
    nrm[30:60] *= rng.rand(30, 1) + 0.5                            # unnormalised shading normals (App. C.4)
    r[:, 18:21] = nrm
    r[:, 21:24] = rng.randn(n, 3)                                  # ray direction (Ray ctor normalises)
    L = rng.randn(n, 3); L /= np.linalg.norm(L, axis=1, keepdims=True)
    r[:, 24:27] = L
    r[:, 27:30] = rng.rand(n, 3)
    return r


def main():
    if not os.path.exists(REF_BIN):
        sys.exit("build oracle/_ref first: make -C oracle")
    tmp = tempfile.mkdtemp(prefix="golden_")
    out = os.path.join(tmp, "o.bin")

    run_ref("xorwow", out)
    np.savez_compressed(os.path.join(HERE, "xorwow.npz"), **read_container(out))
    run_ref("texel", out)
    np.savez_compressed(os.path.join(HERE, "texel.npz"), **read_container(out))

    rec = disney_records()
    rec.tofile(os.path.join(tmp, "disney.in"))
    run_ref("disney", out, os.path.join(tmp, "disney.in"))
    d = read_container(out)
    np.savez_compressed(os.path.join(HERE, "disney.npz"), records=rec, eval_pdf=d["eval_pdf"].reshape(-1, 4), sample=d["sample"].reshape(-1, 3))

    for name, sc in golden_scenes().items():
        flat = os.path.join(tmp, name + ".flat")
        S.save_flat(sc, flat)
        rays = ray_batch(sc)
        rays.tofile(os.path.join(tmp, "rays.in"))
        rng = np.random.RandomState(11)
        rs = np.concatenate([rng.rand(2000), [0.0, 1.0, 0.5, 1e-7, 1 - 1e-7]]).astype(np.float32)
        rs.tofile(os.path.join(tmp, "r.in"))
        dirs = rng.randn(2000, 3); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        dirs = np.concatenate([dirs, np.eye(3), -np.eye(3)]).astype(np.float32)
        dirs.tofile(os.path.join(tmp, "d.in"))
        run_ref("scene", out, flat, os.path.join(tmp, "rays.in"), os.path.join(tmp, "r.in"), os.path.join(tmp, "d.in"))
        d = read_container(out)
        n = len(rays)
        np.savez_compressed(
            os.path.join(HERE, "scene_%s.npz" % name),
            scene_sha=sha(np.frombuffer(open(flat, "rb").read(), np.uint8)),
            bvh_boxes_sha=sha(d["bvh_boxes"]), bvh_meta_sha=sha(d["bvh_meta"]), tri_indices_sha=sha(d["tri_indices"]),
            bvh_boxes_head=d["bvh_boxes"].reshape(-1, 6)[:512], bvh_meta_head=d["bvh_meta"].reshape(-1, 3)[:512],
            cdf_sha=sha(d["cdf"]), cdf_tail=d["cdf"][-16:], radiance_sum=d["radiance_sum"],
            rays=rays, hit_full=d["hit_full"].reshape(n, 14), hit_valid=d["hit_valid"], hit_obj=d["hit_obj"],
            ray_dir_normalised=d["ray_dir_normalised"].reshape(n, 3),
            hdri_r=rs, hdri_xy=d["hdri_xy"].reshape(-1, 2), hdri_dir=d["hdri_dir"].reshape(-1, 3), hdri_pdf=d["hdri_pdf"],
            env_dirs=dirs, env_uv=d["env_uv"].reshape(-1, 2), env_rgb=d["env_rgb"].reshape(-1, 3))
        print(name, "tris", len(sc.tris), "rays", n, "hits", int(d["hit_valid"].sum()))


if __name__ == "__main__":
    main()
