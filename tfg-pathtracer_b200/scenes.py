"""Scene containers, deterministic procedural scene generators and scene-file writers.

The data model is the flat ``ElevenSceneDesc`` of include/eleven_b200.h, i.e. what the reference's
``renderSetup`` copies out of its ``Scene`` (S/kernel.cu:566-661; S/ = reference src/tfg-pathtracer).

Generators:
  cornell_box        config 2: ~36 tris, diffuse Disney, one point light, dim non-black env (F10)
  clock_standin      configs 1/3: ClockCC0 *stand-in* with the published statistics (F1/F2):
                     clock/table/plant = 8265/5184/111832 tris, 3 materials x 4 maps, HDRI, defocus
  displaced_grid     config 4: ~10 M-triangle value-noise height field under a sun+sky HDRI
  textured_lights    config 5: clock stand-in materials + 4 point lights at 3840x2160

Writers: ``save_flat`` / ``load_flat`` (the ELVNSCN1 container, oracle/ref_harness/flat_scene.h) and
``write_reference_scene_dir`` (scene.json + scene.obj + scene.mtl + .bmp + HDRI/*.hdr in the layout
the reference's loader reads), so the same scene can be rendered by the reference.
"""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

TRI_DT = np.dtype([("vertices", "<f4", (3, 3)), ("uv", "<f4", (3, 3)), ("normals", "<f4", (3, 3)),
                   ("tangents", "<f4", (3, 3)), ("tangentsSign", "<f4"), ("objectID", "<i4")])
assert TRI_DT.itemsize == 152
MAT_DT = np.dtype([("albedoTextureID", "<i4"), ("emissionTextureID", "<i4"), ("roughnessTextureID", "<i4"),
                   ("metallicTextureID", "<i4"), ("normalTextureID", "<i4"), ("opacityTextureID", "<i4"),
                   ("albedo", "<f4", 3), ("emission", "<f4", 3), ("opacity", "<f4", 3),
                   ("roughness", "<f4"), ("metallic", "<f4"), ("clearcoatGloss", "<f4"), ("clearcoat", "<f4"),
                   ("anisotropic", "<f4"), ("eta", "<f4"), ("transmission", "<f4"), ("specular", "<f4"),
                   ("specularTint", "<f4"), ("sheenTint", "<f4"), ("subsurface", "<f4"), ("sheen", "<f4")])
assert MAT_DT.itemsize == 108
CAM_DT = np.dtype([("xRes", "<u4"), ("yRes", "<u4"), ("focalLength", "<f4"), ("sensorWidth", "<f4"),
                   ("sensorHeight", "<f4"), ("aperture", "<f4"), ("focusDistance", "<f4"),
                   ("rotation", "<f4", 3), ("position", "<f4", 3), ("bokeh", "<u4")])
assert CAM_DT.itemsize == 56
LIGHT_DT = np.dtype([("position", "<f4", 3), ("radiance", "<f4", 3)])
TEXHDR_DT = np.dtype([("format", "<u4"), ("width", "<i4"), ("height", "<i4"), ("xTile", "<f4"), ("yTile", "<f4"),
                      ("xOffset", "<f4"), ("yOffset", "<f4"), ("filter", "<u4")])

TEX_F32_RGB, TEX_U8_SRGB, TEX_U8_LINEAR, TEX_EXTERNAL = 0, 1, 2, 3


@dataclass
class TextureData:
    """Row 0 is the first row the reference indexes (LDR maps: bottom image row; HDRI: top row)."""
    data: np.ndarray                 # (H, W, 3) float32 or uint8
    format: int = TEX_F32_RGB
    xTile: float = 1.0
    yTile: float = 1.0
    xOffset: float = 0.0
    yOffset: float = 0.0
    filter: int = 0
    path: str = ""

    @property
    def width(self): return int(self.data.shape[1])

    @property
    def height(self): return int(self.data.shape[0])


@dataclass
class SceneData:
    camera: np.ndarray                                   # CAM_DT scalar array (shape ())
    tris: np.ndarray                                     # TRI_DT
    object_material: np.ndarray                          # int32
    materials: np.ndarray                                # MAT_DT
    textures: List[TextureData] = field(default_factory=list)
    hdri: Optional[TextureData] = None
    lights: np.ndarray = field(default_factory=lambda: np.zeros(0, LIGHT_DT))
    name: str = "scene"
    object_names: List[str] = field(default_factory=list)
    material_names: List[str] = field(default_factory=list)

    @property
    def width(self): return int(self.camera["xRes"])

    @property
    def height(self): return int(self.camera["yRes"])


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def make_camera(xres, yres, position=(0, 0, 0), rotation=(0, 0, 0), focal_length=0.035, focus_distance=1e6,
                aperture=2.8, bokeh=False):
    """Camera.hpp:6-40 + SceneLoader.hpp:25-46: sensorWidth fixed 35 mm, sensorHeight = 35 mm * yRes/xRes in float."""
    c = np.zeros((), CAM_DT)
    c["xRes"], c["yRes"] = xres, yres
    c["focalLength"] = np.float32(focal_length)
    sw = np.float32(35 * 0.001)
    c["sensorWidth"] = sw
    c["sensorHeight"] = np.float32(sw * (np.float32(yres) / np.float32(xres)))
    c["aperture"] = np.float32(aperture)
    c["focusDistance"] = np.float32(focus_distance)
    c["rotation"] = np.asarray(rotation, np.float32)
    c["position"] = np.asarray(position, np.float32)
    c["bokeh"] = 1 if bokeh else 0
    return c


def default_material(**kw):
    """Material.hpp:12-36 defaults."""
    m = np.zeros((), MAT_DT)
    for k in ("albedoTextureID", "emissionTextureID", "roughnessTextureID", "metallicTextureID",
              "normalTextureID", "opacityTextureID"):
        m[k] = -1
    m["albedo"] = (0.5, 0.5, 0.5)
    m["opacity"] = (1, 1, 1)
    m["roughness"] = 1.0
    m["specular"] = 0.5
    for k, v in kw.items():
        m[k] = v
    return m


def _normalize(v, eps=1e-20):
    n = np.sqrt((v * v).sum(-1, keepdims=True))
    return v / np.maximum(n, eps)


def tangents_from_uv(P, UV, N):
    """Per-triangle tangent from the UV gradient, Gram-Schmidt against each corner normal.
    (Our own simple generator; the reference's loader runs MikkTSpace, S/ObjLoader.hpp:167-168 —
    parity runs against the reference take the tangents from its scene dump instead.)"""
    e1, e2 = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]
    d1, d2 = UV[:, 1] - UV[:, 0], UV[:, 2] - UV[:, 0]
    det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
    ok = np.abs(det) > 1e-20
    r = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
    T = (e1 * d2[:, 1:2] - e2 * d1[:, 1:2]) * r[:, None]
    Bt = (e2 * d1[:, 0:1] - e1 * d2[:, 0:1]) * r[:, None]
    T = np.where(ok[:, None], T, e1)
    Tc = T[:, None, :] - N * (N * T[:, None, :]).sum(-1, keepdims=True)
    Tc = _normalize(Tc)
    gn = np.cross(e1, e2)
    sign = np.where((np.cross(gn, T) * Bt).sum(-1) < 0, -1.0, 1.0)
    return Tc.astype(np.float32), sign.astype(np.float32)


def make_tris(P, UV, N, object_id, T=None, sign=None):
    """P (n,3,3), UV (n,3,2), N (n,3,3) -> TRI_DT array."""
    n = P.shape[0]
    t = np.zeros(n, TRI_DT)
    t["vertices"] = P.astype(np.float32)
    t["uv"][:, :, :2] = UV.astype(np.float32)
    t["normals"] = N.astype(np.float32)
    if T is None:
        T, sign = tangents_from_uv(P.astype(np.float64), UV.astype(np.float64), N.astype(np.float64))
    t["tangents"] = T
    t["tangentsSign"] = sign if sign is not None else 1.0
    t["objectID"] = object_id
    return t


def grid_mesh(Pg, UVg, Ng):
    """(R, C, 3) vertex grid -> triangles (2*(R-1)*(C-1)); arrays P (n,3,3), UV (n,3,2), N (n,3,3)."""
    R, C = Pg.shape[:2]
    i00 = (np.arange(R - 1)[:, None] * C + np.arange(C - 1)[None, :]).ravel()
    i01, i10, i11 = i00 + 1, i00 + C, i00 + C + 1
    idx = np.concatenate([np.stack([i00, i10, i11], 1), np.stack([i00, i11, i01], 1)], 0)
    Pf, UVf, Nf = Pg.reshape(-1, 3), UVg.reshape(-1, 2), Ng.reshape(-1, 3)
    return Pf[idx], UVf[idx], Nf[idx]


def _flip(P, UV, N):
    return P[:, ::-1].copy(), UV[:, ::-1].copy(), N[:, ::-1].copy()


def lathe(profile_r, profile_y, segments, center=(0, 0, 0), flip=False):
    """Surface of revolution about +Y; profile arrays of length K -> 2*(K-1)*segments triangles, smooth normals."""
    r = np.asarray(profile_r, np.float64)
    y = np.asarray(profile_y, np.float64)
    th = np.linspace(0, 2 * np.pi, segments + 1)
    Pg = np.stack([r[:, None] * np.cos(th)[None, :], np.repeat(y[:, None], segments + 1, 1),
                   r[:, None] * np.sin(th)[None, :]], -1) + np.asarray(center, np.float64)
    dr, dy = np.gradient(r), np.gradient(y)
    nr, ny = dy, -dr                                    # outward normal of the profile curve
    Ng = _normalize(np.stack([nr[:, None] * np.cos(th)[None, :], np.repeat(ny[:, None], segments + 1, 1),
                              nr[:, None] * np.sin(th)[None, :]], -1))
    UVg = np.stack(np.meshgrid(np.linspace(0, 1, segments + 1), np.linspace(0, 1, len(r))), -1)
    P, UV, N = grid_mesh(Pg, UVg, Ng)
    # outward-facing winding: normal should agree with geometric normal
    gn = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
    if ((gn * N[:, 0]).sum(-1) < 0).mean() > 0.5:
        P, UV, N = _flip(P, UV, N)
    if flip:
        P, UV, N = _flip(P, UV, N)
        N = -N
    return P, UV, N


def box_mesh(lo, hi, sub=(1, 1), inward=False):
    """Axis-aligned box with flat normals, each face subdivided sub[0] x sub[1]: 12*sub0*sub1 triangles."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    Ps, UVs, Ns = [], [], []
    for axis in range(3):
        for side in (0, 1):
            a, b = [(1, 2), (2, 0), (0, 1)][axis]
            s, t = np.meshgrid(np.linspace(0, 1, sub[0] + 1), np.linspace(0, 1, sub[1] + 1), indexing="ij")
            Pg = np.zeros(s.shape + (3,))
            Pg[..., axis] = hi[axis] if side else lo[axis]
            Pg[..., a] = lo[a] + s * (hi[a] - lo[a])
            Pg[..., b] = lo[b] + t * (hi[b] - lo[b])
            n = np.zeros(3)
            n[axis] = 1.0 if side else -1.0
            Ng = np.broadcast_to(n, Pg.shape).copy()
            P, UV, N = grid_mesh(Pg, np.stack([s, t], -1), Ng)
            gn = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
            if (gn * N[:, 0]).sum(-1)[0] < 0:
                P, UV, N = _flip(P, UV, N)
            if inward:
                P, UV, N = _flip(P, UV, N)
                N = -N
            Ps.append(P), UVs.append(UV), Ns.append(N)
    return np.concatenate(Ps), np.concatenate(UVs), np.concatenate(Ns)


def quad(p0, p1, p2, p3, normal):
    """Two triangles p0 p1 p2 / p0 p2 p3 with a given flat normal (winding is irrelevant: MT is two-sided)."""
    P = np.array([[p0, p1, p2], [p0, p2, p3]], np.float64)
    UV = np.array([[[0, 0], [1, 0], [1, 1]], [[0, 0], [1, 1], [0, 1]]], np.float64)
    N = np.broadcast_to(np.asarray(normal, np.float64), P.shape).copy()
    return P, UV, N


def disc_fan(center, radius, n, normal=(0, 1, 0)):
    """n triangles fanned around `center` in the plane orthogonal to `normal`."""
    nrm = _normalize(np.asarray(normal, np.float64))
    a = np.cross(nrm, [1.0, 0.0, 0.0])
    if np.linalg.norm(a) < 1e-6:
        a = np.cross(nrm, [0.0, 0.0, 1.0])
    a = _normalize(a)
    b = np.cross(nrm, a)
    th = np.linspace(0, 2 * np.pi, n + 1)
    ring = np.asarray(center, np.float64) + radius * (np.cos(th)[:, None] * a + np.sin(th)[:, None] * b)
    P = np.stack([np.broadcast_to(np.asarray(center, np.float64), (n, 3)), ring[:-1], ring[1:]], 1)
    UV = np.stack([np.full((n, 2), 0.5), 0.5 + 0.5 * np.stack([np.cos(th[:-1]), np.sin(th[:-1])], 1),
                   0.5 + 0.5 * np.stack([np.cos(th[1:]), np.sin(th[1:])], 1)], 1)
    N = np.broadcast_to(nrm, P.shape).copy()
    return P, UV, N


def value_noise(shape, seed, octaves=4, base=8):
    """Seeded multi-octave value noise in [0,1], bilinear-upsampled lattices; float64 (H, W)."""
    H, W = shape
    rng = np.random.RandomState(seed)
    out = np.zeros((H, W))
    amp, tot = 1.0, 0.0
    ys, xs = np.linspace(0, 1, H, endpoint=False), np.linspace(0, 1, W, endpoint=False)
    for o in range(octaves):
        n = base * (2 ** o)
        lat = rng.rand(n + 1, n + 1)
        lat[-1, :] = lat[0, :]
        lat[:, -1] = lat[:, 0]
        fy, fx = ys * n, xs * n
        iy, ix = fy.astype(int), fx.astype(int)
        ty, tx = fy - iy, fx - ix
        ty, tx = ty * ty * (3 - 2 * ty), tx * tx * (3 - 2 * tx)
        # separable smooth interpolation as two small matrix products: out = Wy @ lat @ Wx^T
        Wy = np.zeros((H, n + 1)); Wy[np.arange(H), iy] = 1 - ty; Wy[np.arange(H), iy + 1] += ty
        Wx = np.zeros((W, n + 1)); Wx[np.arange(W), ix] = 1 - tx; Wx[np.arange(W), ix + 1] += tx
        out += amp * ((Wy @ lat) @ Wx.T)
        tot += amp
        amp *= 0.5
    return out / tot


# ------------------------------------------------------------------------------------------------
# Radiance .hdr (RGBE) and 24-bit .bmp codecs (published formats; what the reference reads via stb_image)
# ------------------------------------------------------------------------------------------------
def rgbe_encode(img):
    img = np.asarray(img, np.float32)
    v = img.max(-1)
    m, e = np.frexp(v)
    scale = np.where(v > 1e-32, m * 256.0 / np.where(v > 1e-32, v, 1.0), 0.0)
    rgb = np.clip((img * scale[..., None]).astype(np.int64), 0, 255).astype(np.uint8)
    ee = np.where(v > 1e-32, e + 128, 0).astype(np.uint8)
    return np.concatenate([rgb, ee[..., None]], -1)


def rgbe_decode(rgbe):
    """stb_image's hdr convert: value = byte * 2^(e - 136) in float32, (0,0,0) when e == 0."""
    e = rgbe[..., 3].astype(np.int32)
    f = np.where(e != 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0)).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * f[..., None]).astype(np.float32)


def write_hdr(path, rgbe):
    H, W = rgbe.shape[:2]
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(("-Y %d +X %d\n" % (H, W)).encode())
        if W < 8 or W >= 32768:
            f.write(rgbe.tobytes())
            return
        # new-RLE scanlines made of literal (non-run) chunks only
        nchunk = (W + 127) // 128
        line = bytearray()
        for y in range(H):
            line += bytes([2, 2, W >> 8, W & 255])
            for c in range(4):
                ch = rgbe[y, :, c].tobytes()
                for k in range(nchunk):
                    seg = ch[k * 128:(k + 1) * 128]
                    line.append(len(seg))
                    line += seg
        f.write(bytes(line))


def read_hdr(path):
    with open(path, "rb") as f:
        raw = f.read()
    pos = 0
    def line():
        nonlocal pos
        e = raw.index(b"\n", pos)
        s = raw[pos:e]
        pos = e + 1
        return s
    if line() not in (b"#?RADIANCE", b"#?RGBE"):
        raise ValueError("not a Radiance file")
    while True:
        s = line()
        if s == b"":
            break
    dims = line().split()
    H, W = int(dims[1]), int(dims[3])
    out = np.zeros((H, W, 4), np.uint8)
    buf = np.frombuffer(raw, np.uint8)
    if W < 8 or W >= 32768 or not (buf[pos] == 2 and buf[pos + 1] == 2 and not (buf[pos + 2] & 0x80)):
        return np.frombuffer(raw, np.uint8, H * W * 4, pos).reshape(H, W, 4).copy()
    for y in range(H):
        pos += 4
        for c in range(4):
            x = 0
            while x < W:
                cnt = int(buf[pos]); pos += 1
                if cnt > 128:
                    cnt -= 128
                    out[y, x:x + cnt, c] = buf[pos]; pos += 1
                else:
                    out[y, x:x + cnt, c] = buf[pos:pos + cnt]; pos += cnt
                x += cnt
    return out


def write_bmp(path, rgb_rows_bottom_up):
    """rgb (H, W, 3) uint8 with row 0 = BOTTOM image row (BMP's native order; after stb's vertical flip on
    load this is also row 0 of the reference's Texture::data, S/Texture.hpp:49)."""
    H, W = rgb_rows_bottom_up.shape[:2]
    pad = (-W * 3) % 4
    rows = np.zeros((H, W * 3 + pad), np.uint8)
    rows[:, :W * 3] = rgb_rows_bottom_up[:, :, ::-1].reshape(H, W * 3)
    size = 54 + rows.size
    with open(path, "wb") as f:
        f.write(b"BM" + struct.pack("<IHHI", size, 0, 0, 54))
        f.write(struct.pack("<IiiHHIIiiII", 40, W, H, 1, 24, 0, rows.size, 2835, 2835, 0, 0))
        f.write(rows.tobytes())


def read_bmp(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] != b"BM":
        raise ValueError("not a BMP")
    off = struct.unpack_from("<I", raw, 10)[0]
    W, H, planes, bpp, comp = struct.unpack_from("<iiHHI", raw, 18)
    if bpp != 24 or comp != 0:
        raise ValueError("only uncompressed 24-bit BMP")
    flip = H < 0
    H = abs(H)
    stride = (W * 3 + 3) & ~3
    rows = np.frombuffer(raw, np.uint8, stride * H, off).reshape(H, stride)[:, :W * 3].reshape(H, W, 3)[:, :, ::-1]
    return rows[::-1].copy() if flip else rows.copy()


# ------------------------------------------------------------------------------------------------
# procedural environments
# ------------------------------------------------------------------------------------------------
def constant_env(color, size=(1024, 1024)):
    """HDRI(Vector3 color): 1024x1024 constant texture (S/HDRI.hpp:22-38)."""
    d = np.empty((size[1], size[0], 3), np.float32)
    d[:] = np.asarray(color, np.float32)
    return TextureData(d, TEX_F32_RGB)


def studio_env(width=4096, height=2048, seed=11):
    """Soft boxes + dim fill, max/mean ~ 500; values pass through RGBE so that the
    .hdr file the reference loads decodes to exactly this array."""
    u = (np.arange(width) + 0.5) / width
    v = (np.arange(height) + 0.5) / height
    U, V = np.meshgrid(u, v)
    img = np.full((height, width, 3), 0.02, np.float64) * np.array([0.9, 1.0, 1.2])
    img += 0.06 * value_noise((height, width), seed, 3, 4)[..., None]
    def softbox(cu, cv, su, sv, col):
        du = np.minimum(np.abs(U - cu), 1 - np.abs(U - cu))
        w = 1.0 / (1.0 + np.exp((du - su) * 400.0)) / (1.0 + np.exp((np.abs(V - cv) - sv) * 400.0))
        return w[..., None] * np.asarray(col)
    img += softbox(0.30, 0.30, 0.035, 0.05, (60.0, 56.0, 50.0))
    img += softbox(0.72, 0.36, 0.025, 0.07, (28.0, 30.0, 36.0))
    img += softbox(0.05, 0.18, 0.012, 0.02, (180.0, 170.0, 150.0))
    return TextureData(rgbe_decode(rgbe_encode(img.astype(np.float32))), TEX_F32_RGB)


def sun_sky_env(width=4096, height=2048, seed=4):
    """Procedural sun + sky, sun ~ 5e4 x sky."""
    u = (np.arange(width) + 0.5) / width
    v = (np.arange(height) + 0.5) / height
    U, V = np.meshgrid(u, v)
    sky = (0.25 + 0.75 * np.clip(1 - V * 1.6, 0, 1))[..., None] * np.array([0.35, 0.55, 1.0])
    ground = np.array([0.12, 0.10, 0.08])
    img = np.where((V > 0.5)[..., None], ground, sky)
    d2 = ((U - 0.62) * 2.0) ** 2 + (V - 0.22) ** 2
    img = img + (np.exp(-d2 / (2 * 0.004 ** 2)) * 5e4 * 0.5)[..., None] * np.array([1.0, 0.93, 0.82])
    return TextureData(rgbe_decode(rgbe_encode(img.astype(np.float32))), TEX_F32_RGB)


def material_maps(res, seed, tint):
    """albedo (sRGB), roughness, metallic, normal maps, uint8 (res,res,3) each, from seeded value noise."""
    n1 = value_noise((res, res), seed, 5, 4)
    n2 = value_noise((res, res), seed + 101, 4, 8)
    n3 = value_noise((res, res), seed + 202, 3, 16)
    alb = np.clip((0.35 + 0.6 * n1)[..., None] * np.asarray(tint) + 0.08 * (n2 - 0.5)[..., None], 0, 1)
    rough = np.clip(0.25 + 0.7 * n2, 0, 1)
    metal = np.clip((n3 - 0.55) * 4.0, 0, 1)
    gy, gx = np.gradient(n1 * 0.5 + n2 * 0.5)
    s = res / 64.0
    nrm = _normalize(np.stack([-gx * s * 6, -gy * s * 6, np.ones_like(gx)], -1))
    to8 = lambda a: np.clip(np.round(a * 255.0), 0, 255).astype(np.uint8)
    g3 = lambda a: np.repeat(to8(a)[..., None], 3, -1)
    return to8(alb), g3(rough), g3(metal), to8(nrm * 0.5 + 0.5)


# ------------------------------------------------------------------------------------------------
# scene generators
# ------------------------------------------------------------------------------------------------
def cornell_box(res=1024, env=(0.01, 0.01, 0.01), light=True, env_size=(1024, 1024), tilt=None, box_gap=0.0):
    """Config 2: 5 walls + short box + tall box, Kd-only materials
    (roughness 1, metallic 0 => diffuse Disney), point light (10,10,10) under the ceiling."""
    parts = []
    def add(P, UV, N, obj): parts.append(make_tris(P, UV, N, obj))
    X0, X1, Y0, Y1, Z0, Z1 = -1.0, 1.0, 0.0, 2.0, -1.0, 1.0
    add(*quad((X0, Y0, Z0), (X1, Y0, Z0), (X1, Y0, Z1), (X0, Y0, Z1), (0, 1, 0)), 0)      # floor
    add(*quad((X0, Y1, Z0), (X1, Y1, Z0), (X1, Y1, Z1), (X0, Y1, Z1), (0, -1, 0)), 0)     # ceiling
    add(*quad((X0, Y0, Z1), (X1, Y0, Z1), (X1, Y1, Z1), (X0, Y1, Z1), (0, 0, -1)), 0)     # back wall
    add(*quad((X0, Y0, Z0), (X0, Y0, Z1), (X0, Y1, Z1), (X0, Y1, Z0), (1, 0, 0)), 1)      # left, red
    add(*quad((X1, Y0, Z0), (X1, Y0, Z1), (X1, Y1, Z1), (X1, Y1, Z0), (-1, 0, 0)), 2)     # right, green
    def rot_box(lo, hi, ang, obj):
        P, UV, N = box_mesh(lo, hi)
        c = (np.asarray(lo) + np.asarray(hi)) / 2
        ca, sa = np.cos(ang), np.sin(ang)
        R = np.array([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]])
        add((P - c) @ R.T + c, UV, N @ R.T, obj)
    rot_box((0.15, box_gap, -0.65), (0.75, 0.6, -0.05), 0.3, 3)      # short box
    rot_box((-0.75, box_gap, 0.0), (-0.15, 1.2, 0.6), -0.35, 4)      # tall box
    tris = np.concatenate(parts)
    light_pos = np.array([0.0, 1.8, 0.0])
    if tilt is not None:
        # Rigidly rotate the whole box about its centre so that nothing is axis-aligned.  The reference's slab test
        # turns 0*inf into NaN for rays lying exactly in an axis-aligned box face (S/BVH.hpp:75-99) and
        # then MISSES geometry; an axis-aligned Cornell box under an equirect environment provokes this constantly
        # (pole/equator texels give exactly vertical/horizontal shadow rays along walls/floor).  The parity scenes use
        # the tilted box; the axis-aligned one remains the benchmark scene.
        ax, ay, az = np.deg2rad(tilt)
        Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        Rm = Rz @ Ry @ Rx
        c0 = np.array([0.0, 1.0, 0.0])
        rot = lambda a: ((a.astype(np.float64) - c0) @ Rm.T + c0).astype(np.float32)
        tris["vertices"] = rot(tris["vertices"].reshape(-1, 3)).reshape(-1, 3, 3)
        tris["normals"] = (tris["normals"].reshape(-1, 3).astype(np.float64) @ Rm.T).astype(np.float32).reshape(-1, 3, 3)
        tris["tangents"] = (tris["tangents"].reshape(-1, 3).astype(np.float64) @ Rm.T).astype(np.float32).reshape(-1, 3, 3)
        light_pos = (light_pos - c0) @ Rm.T + c0
    mats = np.stack([default_material(albedo=(0.73, 0.73, 0.73)), default_material(albedo=(0.65, 0.05, 0.05)),
                     default_material(albedo=(0.12, 0.45, 0.15))])
    lights = np.zeros(1 if light else 0, LIGHT_DT)
    if light:
        lights[0]["position"] = light_pos
        lights[0]["radiance"] = (10.0, 10.0, 10.0)
    cam = make_camera(res, res, position=(0.0, 1.0, -3.6), focal_length=0.035)
    return SceneData(cam, tris, np.array([0, 1, 2, 0, 0], np.int32), mats, [], constant_env(env, env_size), lights,
                     "cornell", ["walls", "left", "right", "short_box", "tall_box"], ["white", "red", "green"])


def _pad_with_fan(parts, target, center, radius, normal):
    n = sum(p[0].shape[0] for p in parts)
    rem = target - n
    if rem < 0:
        raise ValueError("mesh already has %d > %d triangles" % (n, target))
    if rem > 0:
        parts.append(disc_fan(center, radius, rem, normal))
    return [np.concatenate([p[i] for p in parts]) for i in range(3)]


def _clock_mesh():
    parts = []
    c = np.array([0.0, 0.86, 0.0])
    # body: squat cylinder on its side approximated by a lathe about Y, then rotated so the face looks at -Z
    prof_r = np.concatenate([np.linspace(0.0, 0.085, 8), np.full(12, 0.09), np.linspace(0.085, 0.0, 8)])
    prof_y = np.concatenate([np.full(8, -0.035), np.linspace(-0.035, 0.035, 12), np.full(8, 0.035)])
    P, UV, N = lathe(prof_r, prof_y, 80)
    Rx = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], float)          # Y axis -> -Z... (face toward camera)
    parts.append((P @ Rx.T + c, UV, N @ Rx.T))
    for sx in (-1, 1):                                                 # two bells (hemispheres)
        ph = np.linspace(0, np.pi / 2, 14)
        Pb, UVb, Nb = lathe(0.035 * np.sin(ph)[::-1], 0.035 * np.cos(ph)[::-1], 40, center=(0, 0, 0))
        parts.append((Pb + c + np.array([sx * 0.055, 0.105, 0.0]), UVb, Nb))
        Pf, UVf, Nf = lathe(np.full(6, 0.008), np.linspace(0, 0.03, 6), 16)   # feet
        parts.append((Pf + c + np.array([sx * 0.05, -0.118, 0.0]), UVf, Nf))
    for ang, ln in ((0.6, 0.06), (2.2, 0.045)):                       # hands
        Ph, UVh, Nh = box_mesh((-0.003, 0.0, -0.001), (0.003, ln, 0.001), (1, 4))
        ca, sa = np.cos(ang), np.sin(ang)
        Rz = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]])
        parts.append((Ph @ Rz.T + c + np.array([0, 0, -0.037]), UVh, Nh @ Rz.T))
    th = np.linspace(0, 2 * np.pi, 49)                                 # rim torus
    tube = np.linspace(0, 2 * np.pi, 13)
    R0, r0 = 0.09, 0.006
    Pg = np.stack([(R0 + r0 * np.cos(tube))[:, None] * np.cos(th)[None, :], (R0 + r0 * np.cos(tube))[:, None] * np.sin(th)[None, :],
                   np.repeat((r0 * np.sin(tube))[:, None], 49, 1)], -1)
    Ng = np.stack([np.cos(tube)[:, None] * np.cos(th)[None, :], np.cos(tube)[:, None] * np.sin(th)[None, :],
                   np.repeat(np.sin(tube)[:, None], 49, 1)], -1)
    UVg = np.stack(np.meshgrid(np.linspace(0, 4, 49), np.linspace(0, 1, 13)), -1)
    Pt, UVt, Nt = grid_mesh(Pg, UVg, Ng)
    parts.append((Pt + c + np.array([0, 0, -0.035]), UVt, Nt))
    return _pad_with_fan(parts, 8265, c + np.array([0, 0, -0.0365]), 0.08, (0, 0, -1))


def _table_mesh():
    # one slab, each face subdivided 18 x 24 -> 6 * 432 quads * 2 = 5184 triangles
    return box_mesh((-0.6, 0.70, -0.35), (0.6, 0.75, 0.35), (18, 24))


def _plant_mesh(seed):
    rng = np.random.RandomState(seed)
    parts = []
    base = np.array([0.33, 0.75, 0.08])
    pr = np.concatenate([np.linspace(0.0, 0.06, 6), np.linspace(0.06, 0.085, 20), np.linspace(0.085, 0.075, 4), np.linspace(0.075, 0.0, 3)])
    py = np.concatenate([np.zeros(6), np.linspace(0.0, 0.14, 20), np.linspace(0.14, 0.145, 4), np.full(3, 0.13)])
    P, UV, N = lathe(pr, py, 60)
    parts.append((P + base, UV, N))
    nleaf, ls, lw = 499, 18, 6
    s = np.linspace(0, 1, ls + 1)
    w = np.linspace(-1, 1, lw + 1)
    S, Wd = np.meshgrid(s, w, indexing="ij")
    for k in range(nleaf):
        az, lean = rng.rand() * 2 * np.pi, 0.25 + 0.9 * rng.rand()
        length, width = 0.16 + 0.22 * rng.rand(), 0.018 + 0.02 * rng.rand()
        droop = 0.4 + 1.4 * rng.rand()
        ang = lean + droop * S * S
        rad = np.cumsum(np.cos(ang[:, :1]) * 0 + np.sin(ang) / ls, 0) * length
        hgt = np.cumsum(np.cos(ang) / ls, 0) * length
        half = width * np.sin(np.pi * np.clip(S, 0.02, 1.0) ** 0.7) * Wd
        curl = 0.25 * width * (Wd ** 2)
        x = rad * np.cos(az) - half * np.sin(az)
        z = rad * np.sin(az) + half * np.cos(az)
        y = hgt + curl + 0.14
        Pg = np.stack([x, y, z], -1) + base + np.array([0.02 * rng.randn(), 0, 0.02 * rng.randn()])
        du, dv = np.gradient(Pg, axis=0), np.gradient(Pg, axis=1)
        Ng = _normalize(np.cross(dv, du))
        UVg = np.stack([S, Wd * 0.5 + 0.5], -1)
        parts.append(grid_mesh(Pg, UVg, Ng))
    return _pad_with_fan(parts, 111832, base + np.array([0, 0.125, 0]), 0.07, (0, 1, 0))


def clock_standin(seed=11, tex_res=4096, xres=1920, yres=1080, env_size=(4096, 2048), lights=0):
    """ClockCC0 STAND-IN (the real asset is not in the reference checkout): same statistics
    as the thesis' console screenshot (F2): clock/table/plant = 8265/5184/111832 triangles, 3 materials
    x {albedo sRGB, roughness, metallic, normal} tex_res^2 8-bit maps, HDRI, strong depth of field."""
    meshes = [_clock_mesh(), _table_mesh(), _plant_mesh(seed)]
    assert [m[0].shape[0] for m in meshes] == [8265, 5184, 111832]
    tris = np.concatenate([make_tris(P, UV * np.array(rep), N, i)
                           for i, ((P, UV, N), rep) in enumerate(zip(meshes, ((1, 1), (3, 2), (1, 1))))])
    textures, mats = [], []
    tints = ((0.9, 0.75, 0.45), (0.55, 0.36, 0.22), (0.25, 0.6, 0.2))
    for i in range(3):
        alb, rough, metal, nrm = material_maps(tex_res, seed * 1000 + i * 17, tints[i])
        if i != 0:
            metal[:] = 0
        b = len(textures)
        # order = the reference loader's std::map iteration: map_Bump, map_Kd, map_Ns, refl (S/SceneLoader.hpp:73)
        textures += [TextureData(nrm, TEX_U8_LINEAR), TextureData(alb, TEX_U8_SRGB),
                     TextureData(rough, TEX_U8_LINEAR), TextureData(metal, TEX_U8_LINEAR)]
        mats.append(default_material(normalTextureID=b, albedoTextureID=b + 1, roughnessTextureID=b + 2, metallicTextureID=b + 3))
    cam_pos = np.array([-0.05, 0.95, -0.62])
    focus = float(np.linalg.norm(np.array([0.0, 0.86, -0.04]) - cam_pos)) - 0.05
    cam = make_camera(xres, yres, position=cam_pos, rotation=(8.0, 12.0, 0.0), focal_length=0.05,
                      focus_distance=focus, aperture=2.8, bokeh=True)
    L = np.zeros(lights, LIGHT_DT)
    for k in range(lights):
        a = 2 * np.pi * k / max(lights, 1)
        L[k]["position"] = (0.9 * np.cos(a), 1.6, 0.9 * np.sin(a) - 0.1)
        L[k]["radiance"] = [(4, 3.6, 3.0), (2.0, 2.4, 3.2), (3, 3, 3), (3.5, 2.5, 2.0)][k % 4]
    return SceneData(cam, tris, np.array([0, 1, 2], np.int32), np.stack(mats), textures,
                     studio_env(env_size[0], env_size[1], seed), L, "clockcc0_standin",
                     ["clock", "table", "plant"], ["clock_mat", "table_mat", "plant_mat"])


def displaced_grid(n=2237, seed=4, xres=1920, yres=1080, env_size=(4096, 2048)):
    """Config 4: n x n vertex grid -> 2(n-1)^2 triangles (n=2237 -> 9 999 392; SURVEY §8d's 3162 would give 19.98 M), 4-octave value-noise height,
    smooth normals, one material, sun+sky HDRI."""
    h = value_noise((n, n), seed, 4, 6).astype(np.float32)
    ext = 10.0
    xs = np.linspace(-ext, ext, n, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    Y = (h * 2.2).astype(np.float32)
    Pg = np.stack([X, Y, Z], -1)
    gx, gz = np.gradient(Y, xs, xs)
    Ng = _normalize(np.stack([-gx, np.ones_like(gx), -gz], -1)).astype(np.float32)
    u = np.linspace(0, 1, n, dtype=np.float32)
    UVg = np.stack(np.meshgrid(u, u, indexing="ij"), -1)
    P, UV, N = grid_mesh(Pg, UVg, Ng)
    T = np.zeros_like(P)
    Tx = _normalize(np.stack([np.ones_like(gx), gx, np.zeros_like(gx)], -1)).astype(np.float32)
    _, _, T = grid_mesh(Pg, UVg, Tx)
    tris = make_tris(P, UV, N, 0, T=T, sign=np.ones(P.shape[0], np.float32))
    cam = make_camera(xres, yres, position=(0.0, 4.5, -11.5), rotation=(18.0, 0.0, 0.0), focal_length=0.03)
    mats = np.stack([default_material(albedo=(0.55, 0.5, 0.42))])
    return SceneData(cam, tris, np.array([0], np.int32), mats, [], sun_sky_env(env_size[0], env_size[1], seed),
                     np.zeros(0, LIGHT_DT), "grid%d" % n, ["terrain"], ["ground"])


def material_zoo(xres=160, yres=90, seed=23, lights=1, env_size=(128, 64)):
    """Parity scene for every shading path the default materials never reach (Disney.hpp:108-253, Texture.hpp:114-142,
    kernel.cu:349,355): the ClockCC0 stand-in geometry with
      clock  clearcoat + clearcoatGloss + anisotropic + sheen + sheenTint + subsurface + specularTint, a FLOAT (F32_RGB),
             bilinear-filtered, non-power-of-two albedo map, a tiled + offset non-power-of-two roughness map, constant metallic
      table  constant emission + an emission-free clearcoat, bilinear 8-bit albedo map of odd size, no normal map
      plant  four congruent 8-bit maps (the packed-record path) + an 8-bit bilinear EMISSION map + sheen / subsurface / anisotropy
    plus a point light, so that all three MIS strategies (environment, point light, emission through the BRDF) carry weight."""
    base = clock_standin(seed=11, tex_res=32, xres=xres, yres=yres, env_size=env_size, lights=lights)
    rng = np.random.RandomState(seed)
    tex = []
    def add(data, fmt, **kw):
        tex.append(TextureData(np.ascontiguousarray(data), fmt, **kw)); return len(tex) - 1
    u8 = lambda shape: rng.randint(0, 256, shape + (3,)).astype(np.uint8)
    smooth = lambda h, w, s_: np.clip(value_noise((h, w), s_, 3, 4)[..., None] * np.array([0.9, 0.7, 0.5]) + 0.1, 0, 1)
    # clock
    a0 = add(smooth(24, 40, seed + 1).astype(np.float32), TEX_F32_RGB, filter=1)
    r0 = add(np.repeat((value_noise((23, 37), seed + 2, 3, 4) * 255).astype(np.uint8)[..., None], 3, -1), TEX_U8_LINEAR, xTile=2.0, yTile=3.0, xOffset=0.25, yOffset=0.4)
    n0 = add(base.textures[0].data, TEX_U8_LINEAR)
    m0 = default_material(albedoTextureID=a0, roughnessTextureID=r0, normalTextureID=n0, metallic=0.35, clearcoat=0.8, clearcoatGloss=0.6,
                          anisotropic=0.5, sheen=0.3, sheenTint=0.4, subsurface=0.3, specularTint=0.5, specular=0.7)
    # table
    a1 = add((smooth(27, 45, seed + 3) * 255).astype(np.uint8), TEX_U8_SRGB, filter=1)
    m1 = default_material(albedoTextureID=a1, roughness=0.45, metallic=0.0, clearcoat=0.3, clearcoatGloss=0.9, emission=(0.8, 0.5, 0.2),
                          specularTint=1.0, sheen=1.0, sheenTint=1.0)
    # plant: packed maps + emission map
    alb, rough, metal, nrm = material_maps(32, seed * 7, (0.25, 0.6, 0.2))
    n2 = add(nrm, TEX_U8_LINEAR); a2 = add(alb, TEX_U8_SRGB); r2 = add(rough, TEX_U8_LINEAR); t2 = add(metal, TEX_U8_LINEAR)
    e2 = add((u8((16, 16)) // 3), TEX_U8_SRGB, filter=1)
    m2 = default_material(normalTextureID=n2, albedoTextureID=a2, roughnessTextureID=r2, metallicTextureID=t2, emissionTextureID=e2,
                          sheen=0.8, sheenTint=0.2, subsurface=0.7, anisotropic=0.8, clearcoat=1.0, clearcoatGloss=0.1, specular=0.3, specularTint=0.3)
    base.materials = np.stack([m0, m1, m2])
    base.textures = tex
    base.name = "material_zoo"
    return base


def tangent_torture(seed=5):
    """Mesh that exercises every branch of the loader's tangent-space pass (host/tangent_space.cpp; the reference: mikktspace through
    S/mikktspaceCallback.hpp): a UV sphere whose seam duplicates vertices with different texture coordinates, one hemisphere with a
    MIRRORED mapping (orientation flips: groups must not cross it), flat-shaded and smooth-shaded parts, triangles with two equal
    positions (degenerate: borrow a neighbour's frame), triangles whose three texture coordinates coincide or are collinear (no
    derivative: "group with anything"), two fans glued along one edge used by four triangles (butterfly), an isolated triangle with a
    degenerate mapping (keeps the default frame), as a second object a plane with rotated / scaled UV islands."""
    rng = np.random.RandomState(seed)
    P, UV, N = [], [], []
    nu, nv = 24, 12
    def sph(i, j):
        th, ph = np.pi * j / nv, 2 * np.pi * i / nu
        p = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
        return 0.5 * p * (1 + 0.15 * np.sin(3 * ph) * np.sin(2 * th)), p
    for j in range(nv):
        for i in range(nu):
            q = [(i, j), (i + 1, j), (i + 1, j + 1), (i, j + 1)]
            pts = [sph(*c) for c in q]
            uvs = [np.array([c[0] / nu, c[1] / nv]) for c in q]
            if i >= nu // 2:
                uvs = [np.array([1.0 - u[0], u[1]]) for u in uvs]                      # mirrored island
            for tri in ((0, 1, 2), (0, 2, 3)):
                pp = np.array([pts[k][0] for k in tri]); nn = np.array([pts[k][1] for k in tri])
                if j in (3, 4):                                                        # a flat-shaded band
                    fn = np.cross(pp[1] - pp[0], pp[2] - pp[0]); fn = fn / (np.linalg.norm(fn) + 1e-30)
                    nn = np.repeat(fn[None], 3, 0)
                P.append(pp); N.append(nn); UV.append(np.array([uvs[k] for k in tri]))
    P, UV, N = [np.array(x) for x in (P, UV, N)]
    n0 = len(P)
    pick = rng.choice(n0, 40, replace=False)
    P[pick[:10], 1] = P[pick[:10], 0]                                                  # two equal positions
    UV[pick[10:20]] = UV[pick[10:20], :1]                                              # all three texture coordinates equal
    UV[pick[20:30], 2] = 0.5 * (UV[pick[20:30], 0] + UV[pick[20:30], 1])               # collinear texture coordinates
    UV[pick[30:40]] = UV[pick[30:40]][:, ::-1]                                         # locally flipped mapping inside an island
    # butterfly: four triangles on one edge
    a, b = np.array([0.0, 1.2, 0.0]), np.array([0.0, 1.6, 0.0])
    wings = [np.array([0.4, 1.4, 0.0]), np.array([-0.4, 1.4, 0.1]), np.array([0.0, 1.4, 0.4]), np.array([0.1, 1.4, -0.4])]
    for k, w in enumerate(wings):
        pp = np.array([a, b, w]) if k % 2 == 0 else np.array([b, a, w])
        fn = np.cross(pp[1] - pp[0], pp[2] - pp[0]); fn /= np.linalg.norm(fn)
        P = np.concatenate([P, pp[None]]); N = np.concatenate([N, np.repeat(fn[None], 3, 0)[None]])
        UV = np.concatenate([UV, np.array([[0.0, 0.0], [0.0, 1.0], [1.0, 0.5]])[None] if k % 2 == 0 else np.array([[0.0, 1.0], [0.0, 0.0], [1.0, 0.5]])[None]])
    # isolated triangle without a mapping
    P = np.concatenate([P, np.array([[[2.0, 0, 0], [2.5, 0, 0], [2.0, 0.5, 0]]])]); N = np.concatenate([N, np.array([[[0, 0, 1.0]] * 3])])
    UV = np.concatenate([UV, np.zeros((1, 3, 2))])
    t0 = make_tris(P, UV, N, 0)
    # second object: plane with rotated / scaled UV islands
    g = 10
    P2, UV2, N2 = [], [], []
    for y in range(g):
        for x in range(g):
            c = np.array([[x, 0, y], [x + 1, 0, y], [x + 1, 0, y + 1], [x, 0, y + 1]], np.float64) / g - np.array([0.5, 0.8, 0.5])
            ang = 0.7 * ((x // 3) + 2 * (y // 3)); s_ = 1.0 + 0.5 * ((x // 3) % 2)
            R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]) * s_
            uv = (np.array([[x, y], [x + 1, y], [x + 1, y + 1], [x, y + 1]], np.float64) / g) @ R.T
            for tri in ((0, 2, 1), (0, 3, 2)):
                P2.append(c[list(tri)]); UV2.append(uv[list(tri)]); N2.append(np.array([[0, 1.0, 0]] * 3))
    t1 = make_tris(np.array(P2), np.array(UV2), np.array(N2), 1)
    base = cornell_box(32)
    base.tris = np.concatenate([t0, t1])
    base.object_material = np.zeros(2, np.int32)
    base.object_names = ["blob", "plane"]
    base.name = "tangent_torture"
    return base


def textured_lights(xres=3840, yres=2160, tex_res=4096, seed=11):
    """Config 5: C1-style textured materials + 4 point lights + defocus + env at 3840x2160."""
    s = clock_standin(seed, tex_res, xres, yres, lights=4)
    s.name = "textured_lights_4k"
    return s


# ------------------------------------------------------------------------------------------------
# flat container I/O (oracle/ref_harness/flat_scene.h)
# ------------------------------------------------------------------------------------------------
def _tex_header(t: TextureData):
    h = np.zeros((), TEXHDR_DT)
    h["format"], h["width"], h["height"] = t.format, t.width, t.height
    h["xTile"], h["yTile"], h["xOffset"], h["yOffset"], h["filter"] = t.xTile, t.yTile, t.xOffset, t.yOffset, t.filter
    return h


def save_flat(scene: SceneData, path):
    with open(path, "wb") as f:
        f.write(b"ELVNSCN1")
        f.write(scene.camera.tobytes())
        for arr in (scene.tris, scene.object_material.astype("<i4"), scene.materials):
            f.write(struct.pack("<I", len(arr)))
            f.write(np.ascontiguousarray(arr).tobytes())
        f.write(struct.pack("<I", len(scene.textures)))
        for t in scene.textures:
            f.write(_tex_header(t).tobytes())
            f.write(np.ascontiguousarray(t.data).tobytes())
        f.write(_tex_header(scene.hdri).tobytes())
        f.write(np.ascontiguousarray(scene.hdri.data, np.float32).tobytes())
        f.write(struct.pack("<I", len(scene.lights)))
        f.write(np.ascontiguousarray(scene.lights).tobytes())


def load_flat(path, texture_loader=None) -> SceneData:
    """Reads an ELVNSCN1 file.  Textures marked external (format 3; the reference harness' --external-textures)
    are loaded from the image files listed in ``<path>.textures.txt`` as 8-bit maps."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != b"ELVNSCN1":
        raise ValueError("bad magic in %s" % path)
    pos = 8
    def take(dt, n=None):
        nonlocal pos
        cnt = 1 if n is None else n
        a = np.frombuffer(raw, dt, cnt, pos).copy()
        pos += dt.itemsize * cnt
        return a[0] if n is None else a
    u32 = np.dtype("<u4")
    cam = take(CAM_DT).reshape(())
    tris = take(TRI_DT, int(take(u32)))
    objm = take(np.dtype("<i4"), int(take(u32)))
    mats = take(MAT_DT, int(take(u32)))
    listing = {}
    if os.path.exists(path + ".textures.txt"):
        for ln in open(path + ".textures.txt"):
            i, cs, p = ln.rstrip("\n").split(" ", 2)
            listing[int(i)] = (cs, p)
    def take_tex(i):
        h = take(TEXHDR_DT)
        W, H, fmt = int(h["width"]), int(h["height"]), int(h["format"])
        if fmt == TEX_EXTERNAL:
            cs, p = listing[i]
            data = (texture_loader or read_bmp)(p)
            fmt = TEX_U8_SRGB if cs == "srgb" else TEX_U8_LINEAR
        elif fmt == TEX_F32_RGB:
            data = take(np.dtype("<f4"), W * H * 3).reshape(H, W, 3)
        else:
            data = take(np.dtype("u1"), W * H * 3).reshape(H, W, 3)
        return TextureData(data, fmt, float(h["xTile"]), float(h["yTile"]), float(h["xOffset"]), float(h["yOffset"]), int(h["filter"]))
    texs = [take_tex(i) for i in range(int(take(u32)))]
    hdri = take_tex(-1)
    lights = take(LIGHT_DT, int(take(u32)))
    return SceneData(cam, tris, objm, mats, texs, hdri, lights, os.path.basename(path))


# ------------------------------------------------------------------------------------------------
# reference-format scene directory
# ------------------------------------------------------------------------------------------------
def write_reference_scene_dir(scene: SceneData, out_dir, env_color=None):
    """Emits scene.json / scene.obj / scene.mtl / textures/*.bmp / HDRI/env.hdr as the reference's loadScene
    reads them (S/SceneLoader.hpp:7-144, S/ObjLoader.hpp:71-171).  OBJ stores -z (the loader negates z,
    S/ObjLoader.hpp:85,91).  Texture paths are absolute because the loader resolves them against the CWD.
    Material scalars other than Kd/Ke cannot be expressed in this format (they stay at Material.hpp defaults)."""
    out_dir = os.path.abspath(out_dir)
    os.makedirs(os.path.join(out_dir, "HDRI"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "textures"), exist_ok=True)
    c = scene.camera
    js = {"camera": {"xRes": int(c["xRes"]), "yRes": int(c["yRes"]),
                     "position": dict(zip("xyz", map(float, c["position"]))),
                     "rotation": dict(zip("xyz", map(float, c["rotation"]))),
                     "focalLength": float(c["focalLength"]), "focusDistance": float(c["focusDistance"]),
                     "aperture": float(c["aperture"]), "bokeh": bool(c["bokeh"])},
          "pointLights": [{"position": dict(zip("xyz", map(float, l["position"]))),
                           "radiance": dict(zip("xyz", map(float, l["radiance"])))} for l in scene.lights]}
    if env_color is not None:
        js["hdri"] = {"color": dict(zip("rgb", map(float, env_color)))}
    else:
        js["hdri"] = {"name": "env"}
        write_hdr(os.path.join(out_dir, "HDRI", "env.hdr"), rgbe_encode(scene.hdri.data))
    with open(os.path.join(out_dir, "scene.json"), "w") as f:
        json.dump(js, f, indent=1)
    tex_paths = []
    for i, t in enumerate(scene.textures):
        p = os.path.join(out_dir, "textures", "tex%02d.bmp" % i)
        if t.format == TEX_F32_RGB:
            raise ValueError("float textures cannot be written as .bmp")
        write_bmp(p, t.data)
        tex_paths.append(p)
    mnames = scene.material_names or ["mat%d" % i for i in range(len(scene.materials))]
    with open(os.path.join(out_dir, "scene.mtl"), "w") as f:
        for m, nm in zip(scene.materials, mnames):
            f.write("newmtl %s\n" % nm)
            f.write("Kd %.8f %.8f %.8f\n" % tuple(m["albedo"]))
            f.write("Ke %.8f %.8f %.8f\n" % tuple(m["emission"]))
            for key, fld in (("map_Kd", "albedoTextureID"), ("map_Ns", "roughnessTextureID"),
                             ("refl", "metallicTextureID"), ("map_Bump", "normalTextureID")):
                if m[fld] >= 0:
                    f.write("%s %s\n" % (key, tex_paths[int(m[fld])]))
            f.write("\n")
    onames = scene.object_names or ["obj%d" % i for i in range(len(scene.object_material))]
    with open(os.path.join(out_dir, "scene.obj"), "w") as f:
        f.write("# generated by tfg-pathtracer_b200/scenes.py\n")
        base = 0
        for oid, nm in enumerate(onames):
            t = scene.tris[scene.tris["objectID"] == oid]
            n = len(t)
            f.write("o %s\nusemtl %s\n" % (nm, mnames[int(scene.object_material[oid])]))
            V = t["vertices"].reshape(-1, 3) * np.array([1, 1, -1], np.float32)
            Nn = t["normals"].reshape(-1, 3) * np.array([1, 1, -1], np.float32)
            UV = t["uv"].reshape(-1, 3)[:, :2]
            f.write("".join("v %.8f %.8f %.8f\n" % tuple(r) for r in V))
            f.write("".join("vt %.8f %.8f\n" % tuple(r) for r in UV))
            f.write("".join("vn %.8f %.8f %.8f\n" % tuple(r) for r in Nn))
            idx = base + 1 + np.arange(3 * n).reshape(n, 3)
            f.write("".join("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a, a, a, b, b, b, cc, cc, cc) for a, b, cc in idx))
            base += 3 * n
    return out_dir


def camera_rays(scene: SceneData, n=None, seed=0):
    """Pinhole rays through pixel centres (no jitter, no lens): n random pixels, or all if n is None.
    Returns float32 (n, 6).  For the closest-hit batches of the parity tests/bench."""
    c = scene.camera
    W, H = int(c["xRes"]), int(c["yRes"])
    if n is None:
        ys, xs = np.divmod(np.arange(W * H), W)
    else:
        rng = np.random.RandomState(seed)
        xs, ys = rng.randint(0, W, n), rng.randint(0, H, n)
    sw, sh, fl = float(c["sensorWidth"]), float(c["sensorHeight"]), float(c["focalLength"])
    d = np.stack([(xs + 0.5) / W * sw - sw / 2, (ys + 0.5) / H * sh - sh / 2, np.full(len(xs), fl)], -1)
    rx, ry, rz = np.deg2rad(np.asarray(c["rotation"], np.float64))
    def rot(d):
        d = np.stack([d[:, 0], d[:, 1] * np.cos(rx) - d[:, 2] * np.sin(rx), d[:, 1] * np.sin(rx) + d[:, 2] * np.cos(rx)], -1)
        d = np.stack([d[:, 0] * np.cos(ry) + d[:, 2] * np.sin(ry), d[:, 1], d[:, 2] * np.cos(ry) - d[:, 0] * np.sin(ry)], -1)
        return np.stack([d[:, 0] * np.cos(rz) - d[:, 1] * np.sin(rz), d[:, 0] * np.sin(rz) + d[:, 1] * np.cos(rz), d[:, 2]], -1)
    d = _normalize(rot(d))
    o = np.broadcast_to(np.asarray(c["position"], np.float64), d.shape)
    return np.concatenate([o, d], -1).astype(np.float32)
