"""ctypes mirror of include/eleven_b200.h and loader of the C-ABI shared library.

The product path has NO fallback: if ``libeleven_b200.so`` is missing or fails to load,
``load_library()`` raises.  (The CPU oracle under oracle/ is test infrastructure and is never
imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import scenes as S

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ELEVEN_LIB") or os.path.join(_HERE, "csrc", "libeleven_b200.so")   # ELEVEN_LIB: A/B-testing kernel variants

PASS_BEAUTY, PASS_DENOISE, PASS_NORMAL, PASS_TANGENT, PASS_BITANGENT = range(5)
RNG_REFERENCE, RNG_FAST = 0, 1
ENV_CDF, ENV_ALIAS = 0, 1
HIT_KEY, HIT_MIN_T = 0, 1
BVH_HOST, BVH_DEVICE = 0, 1
FLAG_TERMINATE_DEAD_PATHS, FLAG_COUNTERS, FLAG_TIME_KERNELS, FLAG_SKIP_NULL_NEE, FLAG_FAST_MATH, FLAG_ANYHIT_LIGHT_SHADOWS = 1, 2, 4, 8, 16, 32
COMM_ID_BYTES = 128


class ElevenConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("rng_mode", C.c_uint32), ("env_mode", C.c_uint32), ("hit_mode", C.c_uint32),
                ("max_bounces", C.c_uint32), ("sample_offset", C.c_uint32), ("sample_stride", C.c_uint32),
                ("flags", C.c_uint32), ("seed", C.c_uint64), ("wave_spp", C.c_uint32), ("bvh_builder", C.c_uint32)]


class ElevenCamera(C.Structure):
    _fields_ = [("xRes", C.c_uint32), ("yRes", C.c_uint32), ("focalLength", C.c_float), ("sensorWidth", C.c_float),
                ("sensorHeight", C.c_float), ("aperture", C.c_float), ("focusDistance", C.c_float),
                ("rotation", C.c_float * 3), ("position", C.c_float * 3), ("bokeh", C.c_uint32)]


class ElevenTexture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("format", C.c_uint32), ("width", C.c_int32), ("height", C.c_int32),
                ("xTile", C.c_float), ("yTile", C.c_float), ("xOffset", C.c_float), ("yOffset", C.c_float),
                ("filter", C.c_uint32)]


class ElevenSceneDesc(C.Structure):
    _fields_ = [("camera", ElevenCamera),
                ("triCount", C.c_uint32), ("tris", C.c_void_p),
                ("objectCount", C.c_uint32), ("objectMaterial", C.c_void_p),
                ("materialCount", C.c_uint32), ("materials", C.c_void_p),
                ("textureCount", C.c_uint32), ("textures", C.POINTER(ElevenTexture)),
                ("hdri", ElevenTexture),
                ("pointLightCount", C.c_uint32), ("pointLights", C.c_void_p)]


class ElevenHit(C.Structure):
    _fields_ = [("tri", C.c_int32), ("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("key", C.c_float)]


# bvh8.h: Node8 (80 B) and TriSlot (48 B) as they live on the device
NODE8_DT = np.dtype([("p", "<f4", 3), ("e", "u1", 3), ("imask", "u1"), ("childBase", "<u4"), ("triBase", "<u4"), ("triMask", "<u4"), ("slack", "<f4"),
                     ("qlox", "u1", 8), ("qloy", "u1", 8), ("qloz", "u1", 8), ("qhix", "u1", 8), ("qhiy", "u1", 8), ("qhiz", "u1", 8)])
SLOT_DT = np.dtype([("v0", "<f4", 3), ("e1", "<f4", 3), ("e2", "<f4", 3), ("tri", "<i4"), ("material", "<i4"), ("shiftBound", "<f4")])
assert NODE8_DT.itemsize == 80 and SLOT_DT.itemsize == 48

HIT_DT = np.dtype([("tri", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("key", "<f4")])


class ElevenStats(C.Structure):
    _fields_ = [("pixel_samples", C.c_uint64), ("rays_extension", C.c_uint64), ("rays_shadow_env", C.c_uint64),
                ("rays_shadow_light", C.c_uint64), ("hit_bounces", C.c_uint64), ("nodes_visited", C.c_uint64),
                ("tris_tested", C.c_uint64), ("kernel_launches", C.c_uint64), ("render_ms", C.c_double),
                ("trace_ms", C.c_double), ("extend_ms", C.c_double), ("shade_ms", C.c_double), ("connect_ms", C.c_double),
                ("other_ms", C.c_double), ("extend_launches", C.c_uint64), ("bvh_build_ms", C.c_double), ("bvh_nodes", C.c_uint32),
                ("bvh_tri_slots", C.c_uint32), ("key_slack", C.c_float), ("samples_done", C.c_uint32), ("key_evals", C.c_uint64),
                ("reduce_ms", C.c_double), ("reduce_calls", C.c_uint64), ("nodes_visited_extend", C.c_uint64), ("tris_tested_extend", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def _tex_struct(t: S.TextureData, keep):
    fmt = t.format
    if fmt == S.TEX_F32_RGB:
        arr = np.ascontiguousarray(t.data, np.float32)
    else:
        arr = np.ascontiguousarray(t.data, np.uint8)
    keep.append(arr)
    return ElevenTexture(arr.ctypes.data, fmt, t.width, t.height, t.xTile, t.yTile, t.xOffset, t.yOffset, t.filter)


def make_scene_desc(scene: S.SceneData):
    """Builds an ElevenSceneDesc over the numpy arrays of `scene`.  Returns (desc, keepalive list)."""
    keep = []
    d = ElevenSceneDesc()
    C.memmove(C.byref(d.camera), scene.camera.tobytes(), C.sizeof(ElevenCamera))
    tris = np.ascontiguousarray(scene.tris)
    objm = np.ascontiguousarray(scene.object_material, np.int32)
    mats = np.ascontiguousarray(scene.materials)
    lights = np.ascontiguousarray(scene.lights)
    keep += [tris, objm, mats, lights]
    d.triCount, d.tris = len(tris), tris.ctypes.data
    d.objectCount, d.objectMaterial = len(objm), objm.ctypes.data
    d.materialCount, d.materials = len(mats), mats.ctypes.data
    texs = (ElevenTexture * max(1, len(scene.textures)))()
    for i, t in enumerate(scene.textures):
        texs[i] = _tex_struct(t, keep)
    keep.append(texs)
    d.textureCount, d.textures = len(scene.textures), C.cast(texs, C.POINTER(ElevenTexture))
    d.hdri = _tex_struct(scene.hdri, keep)
    d.pointLightCount = len(lights)
    d.pointLights = lights.ctypes.data if len(lights) else None
    return d, keep


_lib = None


def load_library(path: str = LIB_PATH):
    """Loads libeleven_b200.so and declares prototypes.  Raises if the extension is missing: there is no
    CPU fallback for the product path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError("CUDA extension not built: %s is missing (run `python -c 'import __graft_entry__ as g; g.build()'`)" % path)
    L = C.CDLL(path)
    vp, sz = C.c_void_p, C.c_size_t
    L.eleven_abi_version.restype = C.c_int
    L.eleven_last_error.restype = C.c_char_p
    L.eleven_init.argtypes = [C.POINTER(ElevenConfig), C.POINTER(vp)]
    L.eleven_destroy.argtypes = [vp]
    L.eleven_destroy.restype = None
    L.eleven_scene_upload.argtypes = [vp, C.POINTER(ElevenSceneDesc)]
    L.eleven_render.argtypes = [vp, C.c_int]
    L.eleven_get_film.argtypes = [vp, C.c_int, vp, sz]
    L.eleven_get_pathcount.argtypes = [vp, vp, sz]
    L.eleven_get_samples.argtypes = [vp]
    L.eleven_get_sample_counts.argtypes = [vp, vp, sz]
    L.eleven_get_stats.argtypes = [vp, C.POINTER(ElevenStats)]
    L.eleven_film_reset.argtypes = [vp]
    L.eleven_set_camera.argtypes = [vp, C.POINTER(ElevenCamera)]
    L.eleven_trace_closest.argtypes = [vp, vp, sz, vp]
    L.eleven_trace_device.argtypes = [vp, vp, sz, vp, C.c_int, C.POINTER(C.c_float)]
    L.eleven_film_sums_device.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    L.eleven_comm_unique_id.argtypes = [vp]
    L.eleven_comm_init_rank.argtypes = [vp, vp, C.c_int, C.c_int]
    L.eleven_comm_init_all.argtypes = [C.POINTER(vp), C.c_int]
    L.eleven_reduce_film.argtypes = [vp, C.c_int, C.c_int]
    L.eleven_get_film_reduced.argtypes = [vp, C.c_int, vp, sz]
    L.eleven_resolve_rgba8_reduced.argtypes = [vp, C.c_int, vp, sz]
    L.eleven_get_sample_counts_reduced.argtypes = [vp, vp, sz]
    L.eleven_test_disney.argtypes = [vp, vp, sz, C.c_int, vp, vp]
    L.eleven_test_hdri.argtypes = [vp, vp, vp, sz, C.c_int, C.c_int, vp, vp, vp]
    L.eleven_test_env_lookup.argtypes = [vp, vp, sz, vp]
    L.eleven_test_hitdata.argtypes = [vp, vp, vp, sz, C.c_int, vp]
    L.eleven_device_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.eleven_device_free.argtypes = [vp, vp]
    L.eleven_device_upload.argtypes = [vp, vp, vp, sz]
    L.eleven_device_download.argtypes = [vp, vp, vp, sz]
    L.eleven_resolve_rgba8.argtypes = [vp, C.c_int, vp, sz]
    L.eleven_bvh_download.argtypes = [vp, vp, sz, vp, sz, vp]
    L.eleven_host_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.eleven_host_free.argtypes = [vp, vp]
    L.eleven_bvh_build_host.argtypes = [vp, C.c_uint32, vp, C.c_int, vp, sz, vp, sz, vp, vp, vp]
    if L.eleven_abi_version() != 3:
        raise RuntimeError("ABI version mismatch")
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "eleven_abi_version", "eleven_last_error", "eleven_init", "eleven_destroy", "eleven_scene_upload",
    "eleven_render", "eleven_get_film", "eleven_get_pathcount", "eleven_get_samples", "eleven_get_sample_counts", "eleven_get_stats",
    "eleven_film_reset", "eleven_set_camera", "eleven_trace_closest", "eleven_trace_device", "eleven_film_sums_device",
    "eleven_comm_unique_id", "eleven_comm_init_rank", "eleven_comm_init_all", "eleven_reduce_film", "eleven_get_film_reduced",
    "eleven_resolve_rgba8_reduced", "eleven_get_sample_counts_reduced", "eleven_test_disney", "eleven_test_hdri",
    "eleven_test_env_lookup", "eleven_test_hitdata", "eleven_device_alloc", "eleven_device_free", "eleven_device_upload",
    "eleven_device_download", "eleven_resolve_rgba8", "eleven_bvh_download", "eleven_host_alloc",
    "eleven_host_free", "eleven_bvh_build_host",
]


def bvh_build_host(tris, tri_material=None, threads=0):
    """(nodes, slots, node_slack, key_slack) from the HOST BVH8 builder alone (eleven_bvh_build_host): no GPU, no context."""
    L = load_library()
    tris = np.ascontiguousarray(tris)
    n = len(tris)
    mat = None if tri_material is None else np.ascontiguousarray(tri_material, np.int32)
    counts = np.zeros(2, np.uint32)
    ks = np.zeros(1, np.float32)
    L.eleven_bvh_build_host(tris.ctypes.data, n, None if mat is None else mat.ctypes.data, threads, None, 0, None, 0, None, counts.ctypes.data, ks.ctypes.data)
    nodes, slots, slack = np.zeros(int(counts[0]), NODE8_DT), np.zeros(int(counts[1]), SLOT_DT), np.zeros(int(counts[0]), np.float32)
    rc = L.eleven_bvh_build_host(tris.ctypes.data, n, None if mat is None else mat.ctypes.data, threads, nodes.ctypes.data, len(nodes),
                                 slots.ctypes.data, len(slots), slack.ctypes.data, counts.ctypes.data, ks.ctypes.data)
    if rc != 0:
        raise RuntimeError("eleven_bvh_build_host: %s" % L.eleven_last_error().decode())
    return nodes, slots, slack, float(ks[0])
