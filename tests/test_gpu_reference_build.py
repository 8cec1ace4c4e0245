"""Point-light scenes against THE REFERENCE'S OWN CUDA BUILD, quantified (BASELINE configs 2 and 5 have a point light).

The reference reads `hdriPdf` uninitialised whenever the environment shadow ray is occluded (S/kernel.cu:344,248,351-353): in the
compiled kernel that is whatever the register last held.  This repo (CUDA path AND oracle) DEFINES it as 0 there, which is
also what makes the three MIS weights a partition of unity.  Round 1 documented the class; this test puts a number on it: the
per-pixel agreement, the difference of the means and the RMSE between our parity render and the reference's render of the same
Cornell box at 256 spp, next to the same three numbers for the SAME box without its light (where the deviation class is inert and
the images must agree to the usual bar)."""
import numpy as np
import pytest

import ref_tools as RT
from gpu_metrics import record
from tfg_pathtracer_b200 import renderer as R
from tfg_pathtracer_b200 import scenes as S

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RT.have_ref("precise"), reason="oracle/_ref not built")]


def compare(light, spp=256, res=256):
    sc = S.cornell_box(res, light=light, tilt=(3.0, 7.0, 2.0), box_gap=0.002)
    ref = RT.run_reference(sc, spp, "precise", env_color=(0.01, 0.01, 0.01))
    r = R.Renderer(**R.PARITY).render_setup(ref.scene); r.render_cuda(spp)
    a, b = r.film()[..., :3], ref.passes["beauty"][..., :3]
    pc = r.get_buffers((R.PASS_BEAUTY,))[1]
    f = R.Renderer(**R.FAST).render_setup(ref.scene); f.render_cuda(spp)
    noise = RT.rmse(f.film()[..., :3], a) / np.sqrt(2.0)        # per-image standard error at this spp, from two independent streams of OUR estimator
    f.close(); r.close()
    return dict(within_tol=RT.within(a, b), mean_ours=float(a.mean()), mean_ref=float(b.mean()), rel_mean_diff=float((a.mean() - b.mean()) / b.mean()),
                rmse=RT.rmse(a, b), noise_sigma=float(noise), rmse_over_sigma=RT.rmse(a, b) / float(noise), pathcount_equal=float((pc == ref.pathcount).mean()))


def test_cornell_without_light_matches_the_reference_build():
    m = record("refbuild_cornell_nolight_256spp", **compare(False))
    assert m["within_tol"] >= 0.995, m
    assert m["pathcount_equal"] >= 0.995, m
    assert abs(m["rel_mean_diff"]) < 1e-3, m


def test_point_light_cornell_against_the_reference_build_is_quantified():
    m = record("refbuild_cornell_pointlight_256spp", **compare(True))
    # Same geometry, same RNG stream, same point-light and BRDF strategies; what differs is the weight given where the environment
    # shadow ray is occluded.  The environment of this scene is 0.01 against a radiance-10 light: the bound below says the images are
    # the same picture; the recorded numbers say how far apart (profiles/r2_test_metrics.jsonl).
    # Measured on B200 (gpurun_out r2a): 41.6 % of the pixels within the parity tolerance, image mean +5.2 % (ours 1.4068, reference 1.3374):
    # with a stale hdriPdf > 0 in the denominator the reference's three MIS weights sum to less than 1 wherever the environment sample is
    # occluded, i.e. it loses light there; ours (and the oracle's) sum to 1.
    assert m["pathcount_equal"] >= 0.995, m                      # the paths themselves are the reference's
    assert 0.0 < m["rel_mean_diff"] < 0.08, m                    # brighter (weights sum to 1), by a few per cent
