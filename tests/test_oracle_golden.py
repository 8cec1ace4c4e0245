"""Pins the CPU oracle (oracle/eleven_oracle.cpp) against vectors produced by THE REFERENCE'S OWN CODE
(tests/golden/make_golden.py -> oracle/_ref/ref_host_vectors, which #includes the reference headers).
Everything that is +,-,*,/,sqrt is compared BIT-EXACTLY; functions containing libm transcendentals are
compared bit-exactly too because both sides run the same glibc on the host (the GPU comparisons in
test_gpu_parity.py carry the tolerances)."""
import hashlib
import os

import numpy as np
import pytest

import make_golden as MG
import oracle_lib as O
from tfg_pathtracer_b200 import scenes as S

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_xorwow_matches_curand():
    g = load("xorwow.npz")
    subs, u, st = g["subsequences"], g["uniforms"].reshape(-1, 32), g["states"].reshape(-1, 6)
    for i, s in enumerate(subs):
        out, state = O.xorwow_uniforms(int(s), 32)
        assert (state == st[i]).all(), "curand_init(0, %d, 0) state" % s
        assert (bits(out) == bits(u[i])).all(), "curand_uniform stream of subsequence %d" % s
    assert ((u > 0) & (u <= 1)).all()


def test_texel_tables_match_stb_fastpow():
    g = load("texel.npz")
    assert (bits(O.texel_table(2.2)) == bits(g["srgb"])).all()
    assert (bits(O.texel_table(1.0)) == bits(g["linear"])).all()
    # the "linear" table is NOT the identity: fastPow drops the low mantissa word (S/stb_image.h:127-136)
    assert (g["linear"] <= np.arange(256, dtype=np.float32) / np.float32(255)).all()


def test_disney_eval_pdf_sample():
    g = load("disney.npz")
    ev, sm = O.disney(g["records"])
    ref_ev, ref_sm = g["eval_pdf"], g["sample"]
    same = bits(ev) == bits(ref_ev)
    both_nan = np.isnan(ev) & np.isnan(ref_ev)
    assert (same | both_nan).all(), "DisneyEval/DisneyPdf differ in %d values" % (~(same | both_nan)).sum()
    same = bits(sm) == bits(ref_sm)
    both_nan = np.isnan(sm) & np.isnan(ref_sm)
    assert (same | both_nan).all(), "DisneySample differs in %d values" % (~(same | both_nan)).sum()
    assert (ref_ev[:, :3] > 0).any() and (ref_ev[:, 3] != 1).any()


@pytest.fixture(scope="module", params=["cornell", "clock", "grid"])
def scene_case(request):
    sc = MG.golden_scenes()[request.param]
    g = load("scene_%s.npz" % request.param)
    tmp = "/tmp/_golden_%s_%d.flat" % (request.param, os.getpid())
    S.save_flat(sc, tmp)
    try:
        assert sha(np.frombuffer(open(tmp, "rb").read(), np.uint8)) == str(g["scene_sha"]), \
            "scene generator drifted from the one the golden vectors were made with"
    finally:
        os.remove(tmp)
    orc = O.Oracle(sc)
    yield sc, g, orc
    orc.close()


def test_reference_bvh_builder(scene_case):
    sc, g, orc = scene_case
    boxes, meta, idx = orc.bvh()
    assert len(boxes) == (2 << 18) - 1
    assert (bits(boxes[:512]) == bits(g["bvh_boxes_head"])).all()
    assert (meta[:512] == g["bvh_meta_head"]).all()
    assert sha(boxes) == str(g["bvh_boxes_sha"])
    assert sha(meta) == str(g["bvh_meta_sha"])
    assert sha(idx) == str(g["tri_indices_sha"])
    assert sorted(idx.tolist()) == list(range(len(sc.tris)))


def test_closest_hit_matches_reference_transverse(scene_case):
    sc, g, orc = scene_case
    hits, full, obj = orc.trace(g["rays"], mode=0, full=True)
    valid = hits["tri"] >= 0
    assert (valid == (g["hit_valid"] != 0)).all()
    assert (obj == g["hit_obj"]).all()
    ref = g["hit_full"]
    m = valid
    assert (bits(full[m]) == bits(ref[m])).all(), "Hit fields differ from BVH::transverse"
    # t is not in the reference's Hit; it is pinned through geomPosition = origin + dir * t for the
    # rays whose hit kept the geometric position (S/Tri.hpp:70,89)
    o, d = g["rays"][:, :3], g["ray_dir_normalised"]
    geo = (o + d * hits["t"][:, None]).astype(np.float32)
    kept = m & (bits(geo) == bits(ref[:, :3])).all(1)
    assert kept.sum() > 0.2 * m.sum()
    assert valid.sum() > 1000


def test_brute_force_agrees_up_to_classified_cases(scene_case):
    sc, g, orc = scene_case
    a = orc.trace(g["rays"], mode=0)
    b = orc.trace(g["rays"], mode=1)
    diff = a["tri"] != b["tri"]
    # allowed: equal-key ties (visit order) and reference slab misses (brute force finds a closer/extra hit)
    tie = diff & (a["tri"] >= 0) & (b["tri"] >= 0) & (a["key"] == b["key"])
    slab_miss = diff & ~tie & ((a["tri"] < 0) | (b["key"] < a["key"]))
    assert (diff == (tie | slab_miss)).all()
    assert slab_miss.sum() <= 0.01 * len(a)


def test_hdri_cdf_sampling_pdf(scene_case):
    sc, g, orc = scene_case
    cdf, rsum = orc.hdri_cdf()
    assert bits(np.float32(rsum)) == bits(g["radiance_sum"])[0]
    assert sha(cdf) == str(g["cdf_sha"])
    xy, d, pdf = orc.hdri_sample(g["hdri_r"])
    assert (xy == g["hdri_xy"]).all()
    assert (bits(d) == bits(g["hdri_dir"])).all()
    same = (bits(pdf) == bits(g["hdri_pdf"])) | (np.isinf(pdf) & np.isinf(g["hdri_pdf"]))
    assert same.all()


def test_environment_lookup(scene_case):
    sc, g, orc = scene_case
    rgb = orc.env_lookup(g["env_dirs"])
    assert (bits(rgb) == bits(g["env_rgb"])).all()
