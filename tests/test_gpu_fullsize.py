"""BASELINE.json's configs at THEIR OWN SIZES, as driver-run assertions (round 1 kept these in builder-run scripts).

  config 1  ClockCC0 stand-in 1920x1080, 16 spp, reference RNG: our parity mode against THE REFERENCE'S OWN CUDA BUILD rendering the
            same scene directory on this box (oracle/_ref/eleven_ref_headless_precise; skipped only where that binary is absent)
  config 3  the same scene at full size: >= 100 k camera + bounce rays against the oracle's REFERENCE TRAVERSAL (BVH::transverse over
            the reference's depth-18 tree, S/BVH.hpp:120-175), bit for bit, brute force only to classify the mismatches
  configs 2, 4, 5  tests/fullsize_config.py's properties P1-P6 (Cornell 1024^2; ~10 M-triangle grid; 3840x2160 + 4 point lights)

Bars: BEAUTY >= 99.5 % of pixels within 1e-3 + 1e-3 |ref|; first-hit AOVs within 1e-5 on >= 99.5 %; path counts equal on >= 99.5 %;
closest hits bit-exact up to the documented classes (equal-key ties, reference slab-test misses).
"""
import os
import sys

import numpy as np
import pytest

import fullsize_config as FC
import make_golden as MG
import oracle_lib as O
import ref_tools as RT
from gpu_metrics import record
from test_gpu_parity import classify_hits
from tfg_pathtracer_b200 import renderer as R
from tfg_pathtracer_b200 import scenes as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clock_full():
    """The benchmark scene (cached on the box by bench.get_scene: the bench of the same round reuses it)."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    flat, _ = bench.get_scene(argparse.Namespace(tex=4096, width=1920, height=1080, workload="clock", grid=0), need_dir=False)
    return S.load_flat(flat)


@pytest.mark.skipif(not RT.have_ref("precise"), reason="oracle/_ref/eleven_ref_headless_precise not built (needs /root/reference at build time)")
def test_config1_full_frame_against_the_reference_cuda_build(clock_full):
    sc = clock_full
    sc.object_names, sc.material_names = ["clock", "table", "plant"], ["clock_mat", "table_mat", "plant_mat"]
    spp = 16
    ref = RT.run_reference(sc, spp, "precise", external_textures=True)
    assert ref.info["width"] == 1920 and ref.info["height"] == 1080 and ref.info["samples_pixel0"] == spp
    r = R.Renderer(**R.PARITY).render_setup(ref.scene)            # the scene as the reference's own loader produced it
    r.render_cuda(spp)
    bufs, pc = r.get_buffers()
    m = record("config1_vs_reference_build",
               beauty_within_tol=RT.within(bufs[R.PASS_BEAUTY][..., :3], ref.passes["beauty"][..., :3]),
               normal_within_1e5=RT.within(bufs[R.PASS_NORMAL][..., :3], ref.passes["normal"][..., :3], 1e-5, 1e-5),
               tangent_within_1e5=RT.within(bufs[R.PASS_TANGENT][..., :3], ref.passes["tangent"][..., :3], 1e-5, 1e-5),
               bitangent_within_1e5=RT.within(bufs[R.PASS_BITANGENT][..., :3], ref.passes["bitangent"][..., :3], 1e-5, 1e-5),
               pathcount_equal=float((pc == ref.pathcount).mean()),
               mean_ours=bufs[R.PASS_BEAUTY][..., :3].mean(), mean_ref=ref.passes["beauty"][..., :3].mean(),
               ref_samples_per_s=ref.info["samples_per_s"], ours_parity_samples_per_s=1920 * 1080 * spp / (r.stats()["render_ms"] * 1e-3))
    assert m["beauty_within_tol"] >= 0.995, m
    assert m["normal_within_1e5"] >= 0.995 and m["tangent_within_1e5"] >= 0.995 and m["bitangent_within_1e5"] >= 0.995, m
    assert m["pathcount_equal"] >= 0.995, m
    assert abs(m["mean_ours"] - m["mean_ref"]) / m["mean_ref"] < 1e-3
    r.close()


def test_config3_fullsize_rays_against_the_reference_traversal(clock_full):
    sc = clock_full
    orc = O.Oracle(sc)                                              # builds the reference's depth-18 binned-SAH tree
    r = R.Renderer(bvh_builder=R.BVH_DEVICE, **R.PARITY).render_setup(sc)
    cam = S.camera_rays(sc, 60000, seed=31).astype(np.float32)
    h0 = r.trace_closest(cam)
    ok = h0["tri"] >= 0
    # bounce rays as the renderer makes them: from the hit point (moved 1 mm along the new direction), random directions
    d = cam[ok, 3:].astype(np.float64); d /= np.linalg.norm(d, axis=1, keepdims=True)
    P = cam[ok, :3] + d * h0["t"][ok, None]
    rng = np.random.RandomState(77)
    w = rng.randn(len(P), 3); w /= np.linalg.norm(w, axis=1, keepdims=True)
    bounce = np.concatenate([P + 1e-3 * w, w], 1).astype(np.float32)
    rays = np.concatenate([cam, bounce, MG.ray_batch(sc, 0, 4096, 4096, seed=12)])
    assert len(rays) >= 100000
    ours = r.trace_closest(rays)
    ref = orc.trace(rays, mode=0, threads=os.cpu_count() or 8)
    same = (ours["tri"] == ref["tri"]) & ((ours["tri"] < 0) | ((ours["t"].view(np.uint32) == ref["t"].view(np.uint32)) & (ours["u"].view(np.uint32) == ref["u"].view(np.uint32)) &
                                                              (ours["v"].view(np.uint32) == ref["v"].view(np.uint32)) & (ours["key"].view(np.uint32) == ref["key"].view(np.uint32))))
    mism = np.nonzero(~same)[0]
    bad = 0
    if len(mism):                                                   # brute force (all 125 281 triangles) only where the two disagree
        brute = orc.trace(rays[mism], mode=1, threads=os.cpu_count() or 8)
        exact, tie, slab, badm = classify_hits(ours[mism], ref[mism], brute)
        bad = int(badm.sum())
        record("config3_fullsize_rays_mismatch_classes", ties=int(tie.sum()), reference_slab_misses=int(slab.sum()), bad=bad)
    m = record("config3_fullsize_rays", rays=len(rays), bit_exact=int(same.sum()), mismatches=int(len(mism)), unclassified=bad, hit_fraction=float((ours["tri"] >= 0).mean()))
    assert bad == 0, m
    assert same.mean() >= 0.995, m
    r.close(); orc.close()


@pytest.mark.parametrize("config,spp,extra", [(2, 256, []), (4, 32, ["--oracle-rays", "1024", "--rays", "524288"]), (5, 16, ["--rays", "524288"])])
def test_fullsize_config_properties(config, spp, extra):
    a = FC.parser().parse_args(["--config", str(config), "--spp", str(spp)] + extra)
    out = FC.run(a)
    record("fullsize_config%d" % config, **out)
    assert out["P1_max_rel_dt"] <= 1e-3 and out["P1_uv_in_range"], out
    assert out["P2_farther_hits_where_target_is_really_hit"] == 0, out
    assert out["P2_aimed_hit_fraction"] >= 0.999, out
    assert out["P3_anyhit_equals_closest_valid"], out
    assert out["P4_within_slack"], out
    assert out["P6_bad"] == 0 and out["P6_bit_exact"] + out["P6_ties"] == a.oracle_rays, out
    assert out["P5_parity_finite"] and out["P5_fast_finite"] and out["P5_parity_dropped_fraction"] < 0.01, out
    assert abs(out["P5_mean_fast"] - out["P5_mean_parity_2spp"]) / out["P5_mean_parity_2spp"] < 0.05, out
    expect = {2: (34, 1024, 1024), 4: (9999392, 1920, 1080), 5: (125281, 3840, 2160)}[config]
    assert (out["tris"], out["width"], out["height"]) == expect
