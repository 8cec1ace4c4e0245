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
