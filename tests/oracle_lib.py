"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tfg_pathtracer_b200 import _capi, scenes as S  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
_lib = None


def build_oracle(force=False):
    src = os.path.join(ROOT, "oracle", "eleven_oracle.cpp")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-pthread", "-fPIC", "-shared",
                               "-Wall", "-Wno-unused-function", src, "-o", ORACLE_SO])
    return ORACLE_SO


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(_capi.ElevenSceneDesc), C.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_bvh_node_count.argtypes = [vp]
        L.orc_bvh_dump.argtypes = [vp, vp, vp, vp]
        L.orc_trace.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, C.c_int, C.c_int]
        L.orc_render.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.orc_render_rows.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_reset.argtypes = [vp]
        L.orc_get_film.argtypes = [vp, C.c_int, vp]
        L.orc_get_counts.argtypes = [vp, vp, vp]
        L.orc_get_ray_counts.argtypes = [vp, vp]
        L.orc_xorwow_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_int, vp, vp]
        L.orc_texel_table.argtypes = [C.c_float, vp]
        L.orc_fastpow.restype = C.c_double
        L.orc_fastpow.argtypes = [C.c_double, C.c_double]
        L.orc_disney_eval_pdf.argtypes = [vp, vp, vp, vp]
        L.orc_disney_sample.argtypes = [vp, vp, vp, vp]
        L.orc_camera_ray.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        L.orc_hdri_sample.argtypes = [vp, vp, C.c_int, vp, vp, vp]
        L.orc_hdri_cdf.argtypes = [vp, vp, vp]
        L.orc_set_env_search.argtypes = [vp, C.c_int]
        L.orc_spherical_mapping.argtypes = [vp, C.c_int, vp]
        L.orc_env_lookup.argtypes = [vp, vp, C.c_int, vp]
        L.orc_hitdata.argtypes = [vp, vp, C.c_int, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The reference's algorithm on the CPU, for one scene."""

    def __init__(self, scene: S.SceneData, build_bvh=True):
        self.scene = scene
        self.L = lib()
        desc, self._keep = _capi.make_scene_desc(scene)
        self.h = self.L.orc_create(C.byref(desc), 1 if build_bvh else 0)
        self.W, self.H = scene.width, scene.height

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bvh(self):
        n = self.L.orc_bvh_node_count(self.h)
        boxes = np.zeros((n, 6), np.float32)
        meta = np.zeros((n, 3), np.int32)
        idx = np.zeros(len(self.scene.tris), np.int32)
        self.L.orc_bvh_dump(self.h, _p(boxes), _p(meta), _p(idx))
        return boxes, meta, idx

    def trace(self, rays, mode=0, full=False, threads=8):
        rays = np.ascontiguousarray(rays, np.float32)
        n = len(rays)
        hits = np.zeros(n, _capi.HIT_DT)
        fu = np.zeros((n, 14), np.float32) if full else None
        ob = np.zeros(n, np.int32) if full else None
        self.L.orc_trace(self.h, _p(rays), n, _p(hits), _p(fu) if full else None, _p(ob) if full else None, mode, threads)
        return (hits, fu, ob) if full else hits

    def render(self, spp, max_bounces=5, threads=8, rows=None):
        if rows is None:
            self.L.orc_render(self.h, spp, max_bounces, threads)
        else:
            self.L.orc_render_rows(self.h, spp, rows[0], rows[1], max_bounces, threads)

    def reset(self):
        self.L.orc_reset(self.h)

    def film(self, p=0):
        out = np.zeros((self.H, self.W, 4), np.float32)
        self.L.orc_get_film(self.h, p, _p(out))
        return out

    def counts(self):
        s = np.zeros(self.W * self.H, np.uint32)
        pc = np.zeros(self.W * self.H, np.uint32)
        self.L.orc_get_counts(self.h, _p(s), _p(pc))
        return s, pc

    def ray_counts(self):
        o = np.zeros(3, np.uint64)
        self.L.orc_get_ray_counts(self.h, _p(o))
        return o

    def hdri_sample(self, r):
        r = np.ascontiguousarray(r, np.float32)
        xy = np.zeros((len(r), 2), np.int32)
        d = np.zeros((len(r), 3), np.float32)
        pdf = np.zeros(len(r), np.float32)
        self.L.orc_hdri_sample(self.h, _p(r), len(r), _p(xy), _p(d), _p(pdf))
        return xy, d, pdf

    def set_env_search(self, exact):
        """False: the reference's binarySearch (default).  True: exact CDF inversion — NOT the reference (whose search returns the
        next texel for half of the uniforms), the yardstick the alias-table mode is compared with."""
        self.L.orc_set_env_search(self.h, 1 if exact else 0)

    def hdri_cdf(self):
        cdf = np.zeros(self.scene.hdri.width * self.scene.hdri.height + 1, np.float32)
        rs = np.zeros(1, np.float32)
        self.L.orc_hdri_cdf(self.h, _p(cdf), _p(rs))
        return cdf, rs[0]

    def env_lookup(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32)
        rgb = np.zeros((len(dirs), 3), np.float32)
        self.L.orc_env_lookup(self.h, _p(dirs), len(dirs), _p(rgb))
        return rgb

    def hitdata(self, attrs, object_ids):
        """generateHitData (S/kernel.cu:54-119) for interpolated attributes (n x 14: position, normal, tangent, bitangent, tu, tv):
        n x 21 floats = HitData scalars, emission, albedo, shading normal."""
        attrs = np.ascontiguousarray(attrs, np.float32).reshape(-1, 14)
        out = np.zeros((len(attrs), 27), np.float32)
        for i in range(len(attrs)):
            self.L.orc_hitdata(self.h, _p(attrs[i]), int(object_ids[i]), _p(out[i]))
        return out[:, :21]

    def camera_ray(self, x, y, r5):
        r5 = np.ascontiguousarray(r5, np.float32)
        out = np.zeros(6, np.float32)
        self.L.orc_camera_ray(self.h, x, y, _p(r5), _p(out))
        return out


def xorwow_uniforms(subsequence, n, seed=0):
    out = np.zeros(n, np.float32)
    st = np.zeros(6, np.uint32)
    lib().orc_xorwow_uniforms(seed, subsequence, n, _p(out), _p(st))
    return out, st


def texel_table(gamma):
    out = np.zeros(256, np.float32)
    lib().orc_texel_table(gamma, _p(out))
    return out


def disney(records):
    """records (n, 30): hd21, rayDir3, L3, r3 -> eval_pdf (n,4), sample (n,3)."""
    records = np.ascontiguousarray(records, np.float32)
    n = len(records)
    ev = np.zeros((n, 4), np.float32)
    sm = np.zeros((n, 3), np.float32)
    L = lib()
    for i in range(n):
        r = records[i]
        L.orc_disney_eval_pdf(_p(r[0:21]), _p(r[21:24]), _p(r[24:27]), _p(ev[i]))
        L.orc_disney_sample(_p(r[0:21]), _p(r[21:24]), _p(r[27:30]), _p(sm[i]))
    return ev, sm
