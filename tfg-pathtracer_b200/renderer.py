"""Host-side mirror of the reference's render interface, over the C ABI (include/eleven_b200.h).

Reference (S/ = src/tfg-pathtracer):          here
  renderSetup(Scene*)        S/kernel.cu:566    Renderer.render_setup(scene)
  renderCuda(Scene*, spp)    S/kernel.cu:665    Renderer.render_cuda(spp)
  getBuffers(RenderData&..)  S/kernel.cu:688    Renderer.get_buffers() -> {pass: (H, W, 4) float32}, pathcount
  getSamples()               S/kernel.cu:712    Renderer.get_samples()

Same argument meaning and data contract (film index W*(H-1-y)+x, RGBA float, A = 1, BEAUTY = mean of per-sample
radiance clamped to [0,10]); errors raise ElevenError with the library's message instead of print-and-continue
(S/kernel.cu:657-659).  There is no CPU fallback: constructing a Renderer without the CUDA library raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from . import scenes as S
from ._capi import (BVH_DEVICE, BVH_HOST, ENV_ALIAS, ENV_CDF, FLAG_COUNTERS, FLAG_TERMINATE_DEAD_PATHS, FLAG_TIME_KERNELS, HIT_KEY, HIT_MIN_T,  # noqa: F401
                    PASS_BEAUTY, PASS_BITANGENT, PASS_NORMAL, PASS_TANGENT, RNG_FAST, RNG_REFERENCE)


class ElevenError(RuntimeError):
    pass


PARITY = dict(rng_mode=RNG_REFERENCE, env_mode=ENV_CDF, hit_mode=HIT_KEY, flags=0)
FAST = dict(rng_mode=RNG_FAST, env_mode=ENV_ALIAS, hit_mode=HIT_KEY, flags=FLAG_TERMINATE_DEAD_PATHS | _capi.FLAG_SKIP_NULL_NEE | _capi.FLAG_FAST_MATH | _capi.FLAG_ANYHIT_LIGHT_SHADOWS,
            bvh_builder=BVH_DEVICE)


class Renderer:
    def __init__(self, device=0, rng_mode=RNG_FAST, env_mode=ENV_ALIAS, hit_mode=HIT_KEY, max_bounces=5,
                 sample_offset=0, sample_stride=1, flags=FLAG_TERMINATE_DEAD_PATHS, seed=0, wave_spp=0,
                 bvh_builder=_capi.BVH_HOST):
        self.L = _capi.load_library()
        if rng_mode == RNG_REFERENCE:
            flags &= ~FLAG_TERMINATE_DEAD_PATHS
        self.cfg = _capi.ElevenConfig(device, rng_mode, env_mode, hit_mode, max_bounces, sample_offset, sample_stride, flags, seed,
                                      wave_spp, bvh_builder)
        self.h = C.c_void_p()
        self._ck(self.L.eleven_init(C.byref(self.cfg), C.byref(self.h)))
        self.W = self.H = 0
        self._keep = None
        self._pinned = []

    def _ck(self, rc):
        if rc != 0:
            raise ElevenError("eleven error %d: %s" % (rc, self.L.eleven_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.L.eleven_host_free(self.h, p)
            self._pinned = []
            self.L.eleven_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- the four reference entry points ------------------------------------------------------------
    def render_setup(self, scene: S.SceneData):
        desc, keep = _capi.make_scene_desc(scene)
        self._ck(self.L.eleven_scene_upload(self.h, C.byref(desc)))
        self.W, self.H = scene.width, scene.height
        return self

    def render_cuda(self, sample_target: int):
        self._ck(self.L.eleven_render(self.h, int(sample_target)))

    def get_buffers(self, passes=(PASS_BEAUTY, PASS_NORMAL, PASS_TANGENT, PASS_BITANGENT)):
        out = {}
        n = self.W * self.H
        for p in passes:
            a = np.empty((self.H, self.W, 4), np.float32)
            self._ck(self.L.eleven_get_film(self.h, p, a.ctypes.data, n))
            out[p] = a
        pc = np.empty(n, np.int32)
        self._ck(self.L.eleven_get_pathcount(self.h, pc.ctypes.data, n))
        return out, pc

    def get_samples(self) -> int:
        v = self.L.eleven_get_samples(self.h)
        if v < 0:
            self._ck(v)
        return v

    def get_sample_counts(self):
        c = np.empty(self.W * self.H, np.uint32)
        self._ck(self.L.eleven_get_sample_counts(self.h, c.ctypes.data, self.W * self.H))
        return c

    # --- extras ---------------------------------------------------------------------------------------
    def film(self, p=PASS_BEAUTY):
        return self.get_buffers((p,))[0][p]

    def set_camera(self, camera):
        cam = _capi.ElevenCamera()
        C.memmove(C.byref(cam), np.asarray(camera).tobytes(), C.sizeof(cam))
        self._ck(self.L.eleven_set_camera(self.h, C.byref(cam)))

    def reset(self):
        self._ck(self.L.eleven_film_reset(self.h))

    def stats(self):
        st = _capi.ElevenStats()
        self._ck(self.L.eleven_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        hits = np.zeros(len(rays), _capi.HIT_DT)
        self._ck(self.L.eleven_trace_closest(self.h, rays.ctypes.data, len(rays), hits.ctypes.data))
        return hits

    def bvh(self):
        """(nodes, slots, node_slack) of the device-resident BVH8 (test hook, eleven_bvh_download)."""
        st = self.stats()
        nodes = np.zeros(st["bvh_nodes"], _capi.NODE8_DT)
        slots = np.zeros(st["bvh_tri_slots"], _capi.SLOT_DT)
        slack = np.zeros(st["bvh_nodes"], np.float32)
        self._ck(self.L.eleven_bvh_download(self.h, nodes.ctypes.data, len(nodes), slots.ctypes.data, len(slots), slack.ctypes.data))
        return nodes, slots, slack

    def resolve_rgba8(self, p=PASS_BEAUTY):
        a = np.empty((self.H, self.W, 4), np.uint8)
        self._ck(self.L.eleven_resolve_rgba8(self.h, p, a.ctypes.data, self.W * self.H))
        return a

    # device-resident ray batches (bench: inputs already in HBM)
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._ck(self.L.eleven_device_alloc(self.h, nbytes, C.byref(p)))
        return p

    def device_free(self, p):
        self._ck(self.L.eleven_device_free(self.h, p))

    def device_upload(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._ck(self.L.eleven_device_upload(self.h, dptr, arr.ctypes.data, arr.nbytes))

    def device_download(self, arr, dptr):
        self._ck(self.L.eleven_device_download(self.h, arr.ctypes.data, dptr, arr.nbytes))

    def trace_device(self, d_rays, n, d_hits, any_hit=False):
        ms = C.c_float()
        self._ck(self.L.eleven_trace_device(self.h, d_rays, n, d_hits, 1 if any_hit else 0, C.byref(ms)))
        return ms.value

    def pinned_array(self, shape, dtype=np.float32):
        """numpy array over page-locked host memory (eleven_host_alloc); freed by close()."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._ck(self.L.eleven_host_alloc(self.h, n, C.byref(p)))
        self._pinned.append(p)
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def film_sums_ptr(self, p=PASS_BEAUTY):
        ptr, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.eleven_film_sums_device(self.h, p, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    # --- multi-GPU: one native ncclReduce of the film records to the root -------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(_capi.COMM_ID_BYTES)
        self._ck(self.L.eleven_comm_unique_id(buf))
        return buf.raw

    def comm_init_rank(self, unique_id: bytes, nranks: int, rank: int):
        assert len(unique_id) == _capi.COMM_ID_BYTES
        self._ck(self.L.eleven_comm_init_rank(self.h, C.c_char_p(unique_id), int(nranks), int(rank)))

    def reduce_film(self, root=0, all_passes=False):
        self._ck(self.L.eleven_reduce_film(self.h, int(root), 1 if all_passes else 0))

    def film_reduced(self, p=PASS_BEAUTY, out=None):
        a = np.empty((self.H, self.W, 4), np.float32) if out is None else out
        self._ck(self.L.eleven_get_film_reduced(self.h, p, a.ctypes.data, self.W * self.H))
        return a

    def resolve_rgba8_reduced(self, p=PASS_BEAUTY):
        a = np.empty((self.H, self.W, 4), np.uint8)
        self._ck(self.L.eleven_resolve_rgba8_reduced(self.h, p, a.ctypes.data, self.W * self.H))
        return a

    def get_sample_counts_reduced(self):
        c = np.empty(self.W * self.H, np.uint32)
        self._ck(self.L.eleven_get_sample_counts_reduced(self.h, c.ctypes.data, self.W * self.H))
        return c

    # --- known-answer hooks for the device shading functions (eleven_test_*) ------------------------------
    def test_disney(self, records, fast_math=False):
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, 30)
        ev, sm = np.zeros((len(rec), 4), np.float32), np.zeros((len(rec), 3), np.float32)
        self._ck(self.L.eleven_test_disney(self.h, rec.ctypes.data, len(rec), int(fast_math), ev.ctypes.data, sm.ctypes.data))
        return ev, sm

    def test_hdri(self, r, r2=None, env_mode=ENV_CDF, fast_math=False):
        r = np.ascontiguousarray(r, np.float32)
        r2 = None if r2 is None else np.ascontiguousarray(r2, np.float32)
        xy, d, pdf = np.zeros((len(r), 2), np.int32), np.zeros((len(r), 3), np.float32), np.zeros(len(r), np.float32)
        self._ck(self.L.eleven_test_hdri(self.h, r.ctypes.data, None if r2 is None else r2.ctypes.data, len(r), int(env_mode), int(fast_math),
                                         xy.ctypes.data, d.ctypes.data, pdf.ctypes.data))
        return xy, d, pdf

    def test_env_lookup(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        rgb = np.zeros((len(dirs), 3), np.float32)
        self._ck(self.L.eleven_test_env_lookup(self.h, dirs.ctypes.data, len(dirs), rgb.ctypes.data))
        return rgb

    def test_hitdata(self, attrs, object_ids, fast_math=False):
        attrs = np.ascontiguousarray(attrs, np.float32).reshape(-1, 14)
        ids = np.ascontiguousarray(object_ids, np.int32)
        out = np.zeros((len(attrs), 21), np.float32)
        self._ck(self.L.eleven_test_hitdata(self.h, attrs.ctypes.data, ids.ctypes.data, len(attrs), int(fast_math), out.ctypes.data))
        return out
