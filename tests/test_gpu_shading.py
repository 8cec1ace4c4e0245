"""GPU tests of the SHADING functions, through the C ABI:

  * known-answer tables: the device DisneyEval/Pdf/Sample, HDRI::sample/pdf and the environment look-up against the golden
    vectors made from THE REFERENCE'S OWN HEADERS (tests/golden/disney.npz, scene_*.npz: S/Disney.hpp:108-253,
    S/HDRI.hpp:130-162, S/Texture.hpp:144-156) — the reference's printBRDFMaterial / printHDRISampling pattern
    (S/kernel.cu:726-794) with the tables compared instead of printed;
  * generateHitData (S/kernel.cu:54-119) on every texture path: bilinear filter, float maps, non-power-of-two sizes, tiling +
    offsets, packed 8-bit records, emission maps — against the oracle;
  * a "material zoo" render: clearcoat, anisotropy, sheen, subsurface, specular tint, emission (MIS strategy 3,
    S/kernel.cu:349,355), a point light — image parity with the oracle under the reference RNG;
  * the alias table against the oracle's CDF (not against our own CDF path).

Tolerances.  FM = false (parity flavour, IEEE div/sqrt, --fmad=false): the device differs from the host only in libm
(sinf/cosf/logf/acosf/atan2f, <= 2 ulp each): values are compared in ulps of the reference value, with an absolute floor for
results that cancel to ~0.  FM = true (production flavour, MUFU approximations, like the reference's -use_fast_math build): relative
tolerances stated at the assertions.
"""
import os

import numpy as np
import pytest

import make_golden as MG
import oracle_lib as O
from gpu_metrics import record
from tfg_pathtracer_b200 import renderer as R
from tfg_pathtracer_b200 import scenes as S

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ulps(a, ref):
    """|a - ref| in units of the float32 spacing at |ref| (inf where exactly one is non-finite, 0 where both are NaN or equal)."""
    a = np.asarray(a, np.float32); ref = np.asarray(ref, np.float32)
    sp = np.spacing(np.maximum(np.abs(ref), np.float32(1e-30)).astype(np.float32))
    with np.errstate(invalid="ignore"):
        d = np.abs(a.astype(np.float64) - ref.astype(np.float64)) / sp
    both_nan = np.isnan(a) & np.isnan(ref)
    same_inf = np.isinf(a) & np.isinf(ref) & (np.sign(a) == np.sign(ref))
    d[both_nan | same_inf] = 0
    d[np.isnan(d)] = np.inf
    return d


@pytest.fixture(scope="module")
def ctx():
    r = R.Renderer(**R.PARITY)
    yield r
    r.close()


def test_disney_tables_parity_flavour(ctx):
    g = np.load(os.path.join(G, "disney.npz"))
    ev, sm = ctx.test_disney(g["records"], fast_math=False)
    ref_ev, ref_sm = g["eval_pdf"], g["sample"]
    # scale of one record: cancellation inside the BRDF sums leaves errors proportional to the largest term, not to the result
    d_ev = ulps(ev, ref_ev); d_sm = ulps(sm, ref_sm)
    scale_ev = np.maximum(np.abs(ref_ev).max(1, keepdims=True), 1e-3)
    abs_ev = np.abs(ev.astype(np.float64) - ref_ev) / scale_ev
    fin = np.isfinite(ref_ev)
    m = record("disney_parity", eval_median_ulp=np.median(d_ev[fin]), eval_p99_ulp=np.percentile(d_ev[fin], 99), eval_max_ulp=d_ev[fin].max(),
               eval_max_rel_to_record_scale=abs_ev[fin].max(), sample_median_ulp=np.median(d_sm), sample_p99_ulp=np.percentile(d_sm, 99),
               sample_max_abs=np.abs(sm - ref_sm).max(), eval_exact_fraction=(d_ev[fin] == 0).mean())
    assert (np.isnan(ev) == np.isnan(ref_ev)).all() and (np.isnan(sm) == np.isnan(ref_sm)).all()
    assert m["eval_p99_ulp"] <= 16 and m["eval_max_rel_to_record_scale"] <= 2e-5          # libm-ulp level
    assert m["sample_p99_ulp"] <= 64 and m["sample_max_abs"] <= 2e-6                       # directions: |component| <= 1
    assert (ref_ev[:, :3] > 0).any() and (ref_ev[:, 3] != 1).any()
    # the branches the default materials never take are in the table: clearcoat, anisotropy, sheen, subsurface, tint
    rec = g["records"]
    for col in (3, 4, 8, 9, 10, 11):
        assert (rec[:, col] > 0).sum() > 500


def test_disney_tables_fast_math_flavour(ctx):
    g = np.load(os.path.join(G, "disney.npz"))
    ev, sm = ctx.test_disney(g["records"], fast_math=True)
    ref_ev, ref_sm = g["eval_pdf"], g["sample"]
    fin = np.isfinite(ref_ev).all(1) & np.isfinite(ev).all(1)
    scale = np.maximum(np.abs(ref_ev).max(1, keepdims=True), 1e-3)
    rel = (np.abs(ev.astype(np.float64) - ref_ev) / scale)[fin]
    m = record("disney_fast", eval_max_rel=rel.max(), eval_p99_rel=np.percentile(rel, 99), sample_max_abs=np.nanmax(np.abs(sm - ref_sm)),
               finite_fraction=fin.mean())
    assert fin.mean() > 0.99
    assert m["eval_p99_rel"] <= 1e-4 and m["eval_max_rel"] <= 5e-3       # MUFU rcp/rsq/lg2/ex2: ~1e-6 each, amplified near grazing angles
    assert m["sample_max_abs"] <= 2e-3                                    # __sinf/__cosf: 2^-21.4 absolute


@pytest.mark.parametrize("name", ["cornell", "clock", "grid"])
def test_hdri_sampling_tables(name):
    sc = MG.golden_scenes()[name]
    g = np.load(os.path.join(G, "scene_%s.npz" % name))
    r = R.Renderer(**R.PARITY).render_setup(sc)
    xy, d, pdf = r.test_hdri(g["hdri_r"], env_mode=R.ENV_CDF)
    assert (xy == g["hdri_xy"]).all(), "HDRI::sample texel (CDF binary search, S/HDRI.hpp:130-162)"
    dd = np.abs(d - g["hdri_dir"]).max()
    fin = np.isfinite(g["hdri_pdf"]) & (g["hdri_pdf"] > 0)
    dp = ulps(pdf[fin], g["hdri_pdf"][fin])
    m = record("hdri_" + name, dir_max_abs=dd, pdf_p99_ulp=np.percentile(dp, 99), pdf_max_ulp=dp.max())
    assert dd <= 1e-6 and m["pdf_max_ulp"] <= 8
    assert (np.isinf(pdf) == np.isinf(g["hdri_pdf"])).all()
    # escaped-ray environment look-up (S/kernel.cu:415-417): same texel except where acos/atan2 land within an ulp of a texel border
    rgb = r.test_env_lookup(g["env_dirs"])
    same = (rgb.view(np.uint32) == g["env_rgb"].view(np.uint32)).all(1)
    record("envlookup_" + name, exact_fraction=same.mean())
    assert same.mean() >= 0.995
    r.close()


def test_alias_table_matches_the_oracle_cdf():
    """North-star item "HDRI importance sampling via an alias table": the texel distribution the table produces is the
    distribution of the reference's CDF (its increments, S/HDRI.hpp:107-128) as the ORACLE builds it — chi-square on 4 M
    draws over all texels, plus exact equality of direction and pdf per chosen texel with the CDF path."""
    sc = MG.golden_scenes()["clock"]
    orc = O.Oracle(sc)
    cdf, _ = orc.hdri_cdf()
    p = np.diff(cdf.astype(np.float64)); p = np.maximum(p, 0); p /= p.sum()
    r = R.Renderer(**R.PARITY).render_setup(sc)
    rng = np.random.RandomState(5)
    n = 1 << 22
    u1, u2 = rng.rand(n).astype(np.float32), rng.rand(n).astype(np.float32)
    xy, d, pdf = r.test_hdri(u1, u2, env_mode=R.ENV_ALIAS)
    W = sc.hdri.width
    idx = xy[:, 1] * W + xy[:, 0]
    cnt = np.bincount(idx, minlength=len(p)).astype(np.float64)
    assert (cnt[p == 0] == 0).all(), "texels of zero probability must never be drawn"
    big = p * n >= 20
    chi2 = (((cnt - p * n) ** 2) / (p * n))[big].sum()
    dof = big.sum() - 1
    z = (chi2 - dof) / np.sqrt(2 * dof)
    m = record("alias_vs_oracle_cdf", chi2=chi2, dof=int(dof), z=z, texels=len(p))
    assert abs(z) < 5, m
    # per texel, direction and pdf are the reference's: the oracle's HDRI::sample maps r -> (texel, direction, pdf); direction and
    # pdf depend on the texel only, so a table texel -> (direction, pdf) is collected from oracle calls (its approximate binary
    # search may answer with a neighbour of the texel a given r falls into, which fills the table just as well)
    q = np.concatenate([((cdf[:-1].astype(np.float64) + cdf[1:]) / 2), cdf[:-1].astype(np.float64), rng.rand(50000)]).astype(np.float32)
    oxy, od, opdf = orc.hdri_sample(q)
    oidx = oxy[:, 1] * W + oxy[:, 0]
    tab_d = np.full((len(p), 3), np.nan, np.float32); tab_p = np.full(len(p), np.nan, np.float32)
    tab_d[oidx] = od; tab_p[oidx] = opdf
    have = ~np.isnan(tab_p[idx])
    assert have.mean() > 0.9
    assert np.abs(d[have] - tab_d[idx[have]]).max() <= 1e-6
    fin = have & np.isfinite(tab_p[idx]) & (tab_p[idx] > 0)
    assert ulps(pdf[fin], tab_p[idx[fin]]).max() <= 8
    assert (np.isinf(pdf[have]) == np.isinf(tab_p[idx[have]])).all()
    r.close(); orc.close()


def test_generate_hit_data_on_every_texture_path():
    sc = S.material_zoo()
    orc = O.Oracle(sc, build_bvh=False)
    r = R.Renderer(**R.PARITY).render_setup(sc)
    rng = np.random.RandomState(9)
    n = 6000
    a = np.zeros((n, 14), np.float32)
    nrm = rng.randn(n, 3); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tan = np.cross(nrm, rng.randn(n, 3)); tan /= np.linalg.norm(tan, axis=1, keepdims=True)
    a[:, 3:6], a[:, 6:9], a[:, 9:12] = nrm, tan, np.cross(nrm, tan)
    a[:, 12:14] = rng.rand(n, 2) * 3.0 - 0.5                    # uv outside [0,1] too: wrap / negative-index clamp
    a[:50, 12:14] = rng.randint(0, 3, (50, 2))                  # exactly on texel / tile borders
    obj = rng.randint(0, 3, n).astype(np.int32)
    ref = orc.hitdata(a, obj)
    for fm, tol in ((False, 2e-6), (True, 2e-3)):
        got = r.test_hitdata(a, obj, fast_math=fm)
        err = np.abs(got.astype(np.float64) - ref) / np.maximum(1.0, np.abs(ref))
        per_obj = {int(o): float(err[obj == o].max()) for o in range(3)}
        record("hitdata_fm%d" % fm, max_err=err.max(), per_object=per_obj)
        # texel choice is integer arithmetic on (int)(u*w): identical; what differs is powf (roughness, metallic) and the normalise
        assert err.max() <= tol, per_obj
        assert (got[:, 12:15].max(0) > 0).all()                 # emission reached (constant and mapped)
    r.close(); orc.close()


@pytest.mark.parametrize("lights", [1, 0])
def test_material_zoo_image_matches_oracle(lights):
    """Clearcoat / GTR1, anisotropy, sheen, subsurface, specular tint, emission through the BRDF strategy (NEE_BRDF_C), bilinear and
    float maps, non-power-of-two textures — rendered with the reference RNG and held against the oracle like the other parity scenes."""
    sc = S.material_zoo(lights=lights)
    spp = 4
    orc = O.Oracle(sc); orc.render(spp)
    r = R.Renderer(**R.PARITY).render_setup(sc); r.render_cuda(spp)
    bufs, pc = r.get_buffers()
    ref = orc.film(0)
    ok = (np.abs(bufs[R.PASS_BEAUTY][..., :3] - ref[..., :3]) <= 1e-3 + 1e-3 * np.abs(ref[..., :3])).all(-1)
    smp, opc = orc.counts()
    m = record("zoo_image_lights%d" % lights, within_tol=ok.mean(), pathcount_equal=(pc.astype(np.uint32) == opc).mean(),
               mean_ours=bufs[R.PASS_BEAUTY][..., :3].mean(), mean_oracle=ref[..., :3].mean())
    assert ok.mean() >= 0.995, m
    assert m["pathcount_equal"] >= 0.995
    for p in (R.PASS_NORMAL, R.PASS_TANGENT, R.PASS_BITANGENT):
        okp = (np.abs(bufs[p][..., :3] - orc.film(p)[..., :3]) <= 1e-5 + 1e-5 * np.abs(orc.film(p)[..., :3])).all(-1)
        assert okp.mean() >= 0.995, (p, okp.mean())
    # the same scene in the production configuration: unbiased w.r.t. a longer oracle render (block means)
    orc.reset(); orc.render(48)
    ref = orc.film(0)[..., :3]
    f = R.Renderer(**R.FAST).render_setup(sc); f.render_cuda(256)
    img = f.film()[..., :3]
    H, W = img.shape[:2]
    blk = lambda x: x[:H // 8 * 8, :W // 8 * 8].reshape(H // 8, 8, W // 8, 8, 3).mean((1, 3))
    rel = np.abs(blk(img) - blk(ref)) / (blk(ref) + 0.02)
    m = record("zoo_fast_lights%d" % lights, mean_fast=img.mean(), mean_oracle=ref.mean(), median_block_rel=np.median(rel), p95_block_rel=np.percentile(rel, 95))
    assert abs(img.mean() - ref.mean()) / ref.mean() < 0.02, m
    assert np.median(rel) < 0.05, m
    f.close(); r.close(); orc.close()
