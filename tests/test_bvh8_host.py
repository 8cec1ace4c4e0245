"""CPU tests of the BVH8 layout contract: the HOST builder's tree (eleven_bvh_build_host: no GPU) is structurally valid
(tests/bvh_check.py) and a CPU model of the kernel's walk (tests/bvh8_walk.py: table-driven child masks, triMask ranks,
FFMA.SAT near planes on the per-ray time scale) finds exactly the hits of a brute-force loop over all triangles."""
import numpy as np
import pytest

import make_golden as MG
from bvh8_walk import EXPAND, PERM, brute_force, walk_closest_t
from bvh_check import validate_bvh8
from tfg_pathtracer_b200 import _capi, scenes as S


def _norm_rays(rays):
    o, d = rays[:, :3].astype(np.float32), rays[:, 3:].astype(np.float32)
    with np.errstate(all="ignore"):
        ln = np.sqrt(((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]).astype(np.float32)).astype(np.float32)
        d = (d / ln[:, None]).astype(np.float32)
    return o, d


def test_mask_tables():
    # perm[o] is the XOR permutation of the 8 child bits (an involution); expand gives every child its 3 triangle bits
    for o in range(8):
        assert sorted(PERM[o][[1 << i for i in range(8)]].tolist()) == [1 << i for i in range(8)]
        assert (PERM[o][PERM[o]] == np.arange(256)).all()
        assert all(PERM[o][1 << i] == 1 << (i ^ o) for i in range(8))
    assert EXPAND[0] == 0 and EXPAND[255] == 0xffffff and EXPAND[0b101] == 0b111000111


@pytest.mark.parametrize("scene", ["cornell", "cornell_axis", "grid", "clock"])
def test_host_tree_valid_and_walk_matches_brute_force(scene):
    if scene == "cornell":
        sc = S.cornell_box(64, env_size=(66, 33), tilt=(3.0, 7.0, 2.0), box_gap=0.002)
    elif scene == "cornell_axis":
        sc = S.cornell_box(64, env_size=(66, 33))                      # axis-aligned: flat boxes, rays inside box faces
    elif scene == "grid":
        sc = S.displaced_grid(n=48, xres=64, yres=36, env_size=(64, 32))
    else:
        sc = S.clock_standin(tex_res=16, xres=64, yres=36, env_size=(64, 32))   # the bench geometry: 125 281 triangles, 13 929 wide nodes
    nodes, slots, slack, key_slack = _capi.bvh_build_host(sc.tris)
    cost, depth, n8 = validate_bvh8(nodes, slots, slack, sc.tris)
    assert n8 == len(nodes) and depth >= 1 and key_slack >= 0
    # same tree from a second build with another thread count (deterministic emission)
    n2, s2, _, _ = _capi.bvh_build_host(sc.tris, threads=1)
    assert n2.tobytes() == nodes.tobytes() and s2.tobytes() == slots.tobytes()
    rays = MG.ray_batch(sc, 96, 96, 64, seed=5) if scene != "clock" else MG.ray_batch(sc, 64, 64, 32, seed=5)
    o, d = _norm_rays(rays)
    visited = [0]
    hits = 0
    for i in range(len(rays)):
        if not np.isfinite(d[i]).all():
            continue
        w = walk_closest_t(nodes, slots, o[i], d[i], visited)
        b = brute_force(slots, o[i], d[i])
        assert w[0] == b[0], "ray %d: walk found triangle %d, brute force %d" % (i, w[0], b[0])
        if b[0] >= 0:
            hits += 1
            assert (np.array(w[1:], np.float32).view(np.uint32) == np.array(b[1:], np.float32).view(np.uint32)).all()
    assert hits > len(rays) // 4
    assert visited[0] / len(rays) < 60        # the walk culls (a full sweep would visit every node for every ray)


def test_host_builder_hook_edge_cases_and_errors():
    sc = S.cornell_box(64, env_size=(66, 33))
    nodes, slots, slack, ks = _capi.bvh_build_host(sc.tris[:0])                 # empty scene: one empty root, no slots
    assert len(nodes) == 1 and len(slots) == 0 and nodes["imask"][0] == 0 and nodes["triMask"][0] == 0
    nodes, slots, slack, ks = _capi.bvh_build_host(sc.tris[:1])                 # one triangle: a root with one leaf child
    assert len(nodes) == 1 and len(slots) == 1 and nodes["triMask"][0] == 1 and nodes["imask"][0] == 0
    validate_bvh8(nodes, slots, slack, sc.tris[:1])
    L = _capi.load_library()
    counts = np.zeros(2, np.uint32)
    small = np.zeros(1, _capi.NODE8_DT)
    t = np.ascontiguousarray(sc.tris)
    rc = L.eleven_bvh_build_host(t.ctypes.data, len(t), None, 1, small.ctypes.data, 1, None, 0, None, counts.ctypes.data, None)
    assert rc == -1 and b"too small" in L.eleven_last_error() and counts[0] > 1 and counts[1] == len(t)   # sizes come back with the error
    assert L.eleven_bvh_build_host(None, 5, None, 1, None, 0, None, 0, None, counts.ctypes.data, None) == -1
    assert b"null" in L.eleven_last_error()


def test_presplit_separates_the_wedges_of_a_fan(monkeypatch):
    """What triangle pre-splitting is for (bvh8_build.cpp: presplitTriangles): a disc modelled as a fan of sliver wedges, all of whose
    boxes cover most of the disc.  Rays through the disc must find the same hits through either tree, with far fewer triangle tests
    through the pre-split one, and the pre-split tree is valid (bvh_check: point location on every split triangle)."""
    import bvh8_walk as BW
    nw, R0 = 96, 0.5
    ang = np.linspace(0, 2 * np.pi, nw + 1)
    P = np.zeros((nw, 3, 3)); P[:, 1, 0] = R0 * np.cos(ang[:-1]); P[:, 1, 1] = R0 * np.sin(ang[:-1]); P[:, 2, 0] = R0 * np.cos(ang[1:]); P[:, 2, 1] = R0 * np.sin(ang[1:])
    g = 24                                                             # a finely tessellated floor behind it keeps the scene's mean box small
    xs = np.linspace(-1, 1, g + 1)
    Q = []
    for i in range(g):
        for j in range(g):
            a, b, c, d = (xs[i], xs[j], -1.0), (xs[i + 1], xs[j], -1.0), (xs[i + 1], xs[j + 1], -1.0), (xs[i], xs[j + 1], -1.0)
            Q += [[a, b, c], [a, c, d]]
    V = np.concatenate([P, np.array(Q)])
    tris = S.make_tris(V, np.zeros((len(V), 3, 2)), np.tile(np.array([0, 0, 1.0]), (len(V), 3, 1)), 0)
    rng = np.random.RandomState(3)
    rad, phi = R0 * np.sqrt(rng.rand(150)) * 0.98, 2 * np.pi * rng.rand(150)
    tgt = np.stack([rad * np.cos(phi), rad * np.sin(phi), np.zeros(150)], 1)
    org = tgt + np.array([0.3, 0.2, 2.0]) + 0.05 * rng.normal(size=(150, 3))
    o, d = _norm_rays(np.concatenate([org, tgt - org], 1))
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ELEVEN_PRESPLIT", flag)
        nodes, slots, slack, _ = _capi.bvh_build_host(tris)
        validate_bvh8(nodes, slots, slack, tris)
        calls = [0]
        orig = BW.moller_trumbore

        def counting(o_, d_, v0, e1, e2):
            calls[0] += len(v0)
            return orig(o_, d_, v0, e1, e2)
        monkeypatch.setattr(BW, "moller_trumbore", counting)
        hits = [BW.walk_closest_t(nodes, slots, o[i], d[i]) for i in range(len(o))]
        monkeypatch.setattr(BW, "moller_trumbore", orig)
        res[flag] = (hits, calls[0] / len(o), len(slots))
    print("fan: triangle tests per ray %.1f unsplit, %.1f pre-split; slots %d -> %d" % (res["0"][1], res["1"][1], res["0"][2], res["1"][2]))
    assert res["0"][2] == len(tris) and res["1"][2] > len(tris)            # the wedges own several slots each
    for a, b in zip(res["0"][0], res["1"][0]):
        assert a[0] == b[0] and np.float32(a[1]).view(np.uint32) == np.float32(b[1]).view(np.uint32)
    assert all(h[0] >= 0 and h[0] < nw for h in res["1"][0])               # every ray hits the disc
    assert res["1"][1] < 0.6 * res["0"][1], (res["0"][1], res["1"][1])   # 13.2 -> 6.3 tests per ray here (the clock face of the benchmark scene: 80-99 -> ~12 for rays through it)
