"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star):
  * closest hit (triangle id, t, u, v, key): BIT-EXACT against the oracle's reference traversal, up to the documented
    classes: equal-key ties (visit order) and reference slab-test misses (where the oracle's brute force agrees with us);
  * image (rng = reference XORWOW, env = reference CDF search): >= 99.5 % of pixels within 1e-3 + 1e-3*|ref| on linear
    BEAUTY (outliers = paths that diverged on a libm-ulp difference), first-hit AOVs to 1e-5, path counts equal on
    >= 99.5 % of pixels;
  * fast mode (counter RNG + alias table + dead-path termination): unbiased w.r.t. the oracle.
"""
import numpy as np
import pytest

import bvh_check
import make_golden as MG
import oracle_lib as O
from tfg_pathtracer_b200 import renderer as R
from tfg_pathtracer_b200 import scenes as S

pytestmark = pytest.mark.gpu

ABS_TOL, REL_TOL, PIXEL_FRACTION = 1e-3, 1e-3, 0.995


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def scenes():
    return MG.golden_scenes()


def classify_hits(ours, ref, brute):
    """Returns (exact, tie, slab_miss, bad) boolean masks."""
    same_tri = ours["tri"] == ref["tri"]
    exact = same_tri & ((ours["tri"] < 0) | ((bits(ours["t"]) == bits(ref["t"])) & (bits(ours["u"]) == bits(ref["u"])) &
                                             (bits(ours["v"]) == bits(ref["v"])) & (bits(ours["key"]) == bits(ref["key"]))))
    tie = ~exact & (ours["tri"] >= 0) & (ref["tri"] >= 0) & (bits(ours["key"]) == bits(ref["key"]))
    same_as_brute = (ours["tri"] == brute["tri"]) & ((ours["tri"] < 0) | (bits(ours["t"]) == bits(brute["t"])))
    brute_tie = (ours["tri"] >= 0) & (brute["tri"] >= 0) & (bits(ours["key"]) == bits(brute["key"]))
    slab_miss = ~exact & ~tie & (same_as_brute | brute_tie)
    bad = ~(exact | tie | slab_miss)
    return exact, tie, slab_miss, bad


@pytest.mark.parametrize("name", ["cornell", "clock", "grid"])
def test_closest_hit_bit_exact(scenes, name):
    sc = scenes[name]
    orc = O.Oracle(sc)
    rays = np.concatenate([MG.ray_batch(sc, 4096, 4096, 2048, seed=21), MG.ray_batch(sc)])
    ref = orc.trace(rays, mode=0)
    brute = orc.trace(rays, mode=1)
    r = R.Renderer(**R.PARITY).render_setup(sc)
    ours = r.trace_closest(rays)
    exact, tie, slab, bad = classify_hits(ours, ref, brute)
    assert bad.sum() == 0, "unclassified mismatches at rays %s" % np.nonzero(bad)[0][:10]
    assert exact.mean() > 0.98, (exact.mean(), tie.sum(), slab.sum())
    assert (ours["tri"] >= 0).sum() > 1000
    r.close(); orc.close()


@pytest.mark.parametrize("name", ["cornell", "clock", "grid"])
def test_bvh8_of_both_builders_is_valid_and_device_build_is_hit_exact(scenes, name):
    """the BVH8 built on the GPU.  Both trees are validated on the host (bvh_check), the device
    tree's SAH cost is the host builder's (same algorithm, different leaf policy: within 10 %), it is deterministic, and
    closest hits through it are the oracle's bit for bit (the result does not depend on the tree)."""
    sc = scenes[name]
    host = R.Renderer(**R.PARITY).render_setup(sc)
    dev = R.Renderer(bvh_builder=R.BVH_DEVICE, **R.PARITY).render_setup(sc)
    ch, dh, nh = bvh_check.validate_bvh8(*host.bvh(), sc.tris)
    cd, dd, nd = bvh_check.validate_bvh8(*dev.bvh(), sc.tris)
    assert cd <= 1.10 * ch, (cd, ch)
    assert dd < 40
    dev2 = R.Renderer(bvh_builder=R.BVH_DEVICE, **R.PARITY).render_setup(sc)
    for a, b in zip(dev.bvh(), dev2.bvh()):
        assert a.tobytes() == b.tobytes(), "device build must be deterministic"
    assert abs(dev.stats()["key_slack"] - host.stats()["key_slack"]) <= 1e-6 * host.stats()["key_slack"]
    # the tree the CPU suite checks (eleven_bvh_build_host + tests/bvh8_walk.py) IS the tree the kernels walk
    from tfg_pathtracer_b200 import _capi
    hn, hs, hk, _ = _capi.bvh_build_host(sc.tris, np.asarray(sc.object_material, np.int32)[sc.tris["objectID"]])
    for a, b in zip(host.bvh(), (hn, hs, hk)):
        assert a.tobytes() == b.tobytes(), "host-only builder hook and the uploaded host-built tree differ"
    orc = O.Oracle(sc)
    rays = np.concatenate([MG.ray_batch(sc, 4096, 4096, 2048, seed=5), MG.ray_batch(sc)])
    ref, brute = orc.trace(rays, mode=0), orc.trace(rays, mode=1)
    exact, tie, slab, bad = classify_hits(dev.trace_closest(rays), ref, brute)
    assert bad.sum() == 0 and exact.mean() > 0.98
    oh = host.trace_closest(rays)
    od = dev.trace_closest(rays)
    assert oh.tobytes() == od.tobytes(), "closest hits must not depend on the builder"
    # and a render through the device-built tree equals the render through the host-built one, bit for bit
    host.render_cuda(2); dev.render_cuda(2)
    assert (bits(host.film()) == bits(dev.film())).all()
    for r in (host, dev, dev2):
        r.close()
    orc.close()


def test_device_bvh_chunked_binning_and_overflow_retry(scenes, monkeypatch):
    """The device builder bins the active nodes of a level in chunks (bounded scratch) and sizes the wide-node buffers for
    n/2 nodes, rebuilding with worst-case buffers on overflow.  Both paths must give the very same tree."""
    sc = scenes["clock"]
    ref = R.Renderer(bvh_builder=R.BVH_DEVICE, **R.PARITY).render_setup(sc)
    want = [a.tobytes() for a in ref.bvh()]
    ref.close()
    for env in ({"ELEVEN_BVH_TEST_BIN_CHUNK": "97"}, {"ELEVEN_BVH_TEST_WIDE_DIV": "64"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = R.Renderer(bvh_builder=R.BVH_DEVICE, **R.PARITY).render_setup(sc)
        assert [a.tobytes() for a in r.bvh()] == want, env
        r.close()
        for k in env:
            monkeypatch.delenv(k)


def test_device_bvh_degenerate_inputs():
    """Duplicated triangles (all centroids coincide: median split), a single triangle, 4 triangles."""
    base = S.cornell_box(16, env_size=(8, 8))
    for tris in (np.repeat(base.tris[:1], 37), base.tris[:1].copy(), base.tris[:4].copy(), np.concatenate([base.tris, base.tris, base.tris])):
        sc = S.cornell_box(16, env_size=(8, 8))
        sc.tris = np.ascontiguousarray(tris)
        dev = R.Renderer(bvh_builder=R.BVH_DEVICE, **R.PARITY).render_setup(sc)
        bvh_check.validate_bvh8(*dev.bvh(), sc.tris)
        host = R.Renderer(**R.PARITY).render_setup(sc)
        rays = MG.ray_batch(base, 512, 512, 256, seed=2)
        assert host.trace_closest(rays).tobytes() == dev.trace_closest(rays).tobytes()
        dev.close(); host.close()


@pytest.mark.parametrize("name", ["cornell", "clock"])
def test_min_t_mode_differs_only_inside_key_slack(scenes, name):
    sc = scenes[name]
    rays = MG.ray_batch(sc, 4096, 2048, 1024, seed=5)
    a = R.Renderer(**R.PARITY).render_setup(sc)
    cfg = dict(R.PARITY); cfg["hit_mode"] = R.HIT_MIN_T
    b = R.Renderer(**cfg).render_setup(sc)
    ha, hb = a.trace_closest(rays), b.trace_closest(rays)
    assert ((ha["tri"] >= 0) == (hb["tri"] >= 0)).all()
    d = ha["tri"] != hb["tri"]
    slack = a.stats()["key_slack"]
    assert (np.abs(ha["t"][d] - hb["t"][d]) <= 2 * slack + 1e-5).all()
    a.close(); b.close()


def render_pair(sc, spp):
    orc = O.Oracle(sc)
    orc.render(spp)
    r = R.Renderer(**R.PARITY).render_setup(sc)
    r.render_cuda(spp)
    return orc, r


@pytest.mark.parametrize("name,spp", [("cornell", 8), ("clock", 4), ("grid", 4)])
def test_image_matches_oracle_with_reference_rng(scenes, name, spp):
    sc = scenes[name]
    orc, r = render_pair(sc, spp)
    bufs, pc = r.get_buffers()
    ref = orc.film(0)
    img = bufs[R.PASS_BEAUTY]
    ok = (np.abs(img[..., :3] - ref[..., :3]) <= ABS_TOL + REL_TOL * np.abs(ref[..., :3])).all(-1)
    assert ok.mean() >= PIXEL_FRACTION, "only %.4f of pixels within tolerance" % ok.mean()
    assert (img[..., 3] == 1).all()
    for p in (R.PASS_NORMAL, R.PASS_TANGENT, R.PASS_BITANGENT):
        okp = (np.abs(bufs[p][..., :3] - orc.film(p)[..., :3]) <= 1e-5 + 1e-5 * np.abs(orc.film(p)[..., :3])).all(-1)
        assert okp.mean() >= PIXEL_FRACTION, (p, okp.mean())
    smp, opc = orc.counts()
    assert (pc.astype(np.uint32) == opc).mean() >= PIXEL_FRACTION
    assert r.get_samples() == int(smp[0])
    assert (r.get_sample_counts() == smp).mean() >= PIXEL_FRACTION
    st = r.stats()
    oc = orc.ray_counts()
    assert abs(int(st["rays_extension"]) - int(oc[0])) <= 0.002 * int(oc[0])
    assert st["hit_bounces"] == int(pc.sum())
    r.close(); orc.close()


def test_progressive_render_equals_one_shot(scenes):
    sc = scenes["cornell"]
    a = R.Renderer(**R.PARITY).render_setup(sc); a.render_cuda(6)
    b = R.Renderer(**R.PARITY).render_setup(sc); b.render_cuda(2); b.render_cuda(4)
    assert (bits(a.film()) == bits(b.film())).all()
    b.reset(); b.render_cuda(6)
    assert (bits(a.film()) == bits(b.film())).all()
    a.close(); b.close()


def test_fast_mode_is_unbiased_and_split_invariant(scenes):
    sc = scenes["cornell"]
    orc = O.Oracle(sc); orc.render(64)
    ref = orc.film(0)[..., :3]
    r = R.Renderer(**R.FAST).render_setup(sc); r.render_cuda(256)
    img = r.film()[..., :3]
    assert abs(img.mean() - ref.mean()) / ref.mean() < 0.02
    # blocks of 8x8 pixels average out the noise
    H, W = img.shape[:2]
    blk = lambda a: a[:H // 8 * 8, :W // 8 * 8].reshape(H // 8, 8, W // 8, 8, 3).mean((1, 3))
    rel = np.abs(blk(img) - blk(ref)) / (blk(ref) + 0.02)
    assert np.median(rel) < 0.05
    # sample split: 2 contexts rendering the even/odd samples sum to the same film as 1 context
    one = R.Renderer(**R.FAST).render_setup(sc); one.render_cuda(8)
    parts = []
    for g in range(2):
        cfg = dict(R.FAST); cfg.update(sample_offset=g, sample_stride=2)
        p = R.Renderer(**cfg).render_setup(sc); p.render_cuda(4)
        cnt = p.get_sample_counts().reshape(p.H, p.W, 1)
        parts.append((p.film()[..., :3] * cnt, cnt)); p.close()
    total = parts[0][1] + parts[1][1]
    assert (total.ravel() == one.get_sample_counts()).all()
    np.testing.assert_allclose((parts[0][0] + parts[1][0]) / np.maximum(total, 1), one.film()[..., :3], rtol=2e-5, atol=1e-6)
    one.close(); r.close(); orc.close()


def test_wave_spp_does_not_change_the_film(scenes):
    """K samples of every pixel in flight per wave (ElevenConfig.wave_spp): the counter-based RNG is keyed by the global
    sample index and the K samples of a pixel are accumulated in sample order, so the film is bit-identical for any K,
    including ragged tails (11 = 8 + 2 + 1) and progressive calls."""
    for name in ("cornell", "clock"):
        sc = scenes[name]
        films, counts, stats = [], [], []
        for k in (1, 4, 8, 0):
            cfg = dict(R.FAST); cfg.update(wave_spp=k, sample_offset=3, sample_stride=2)
            r = R.Renderer(**cfg).render_setup(sc)
            if k == 4:
                r.render_cuda(5); r.render_cuda(6)
            else:
                r.render_cuda(11)
            b, pc = r.get_buffers()
            films.append(b); counts.append((pc, r.get_sample_counts())); stats.append(r.stats()); r.close()
        for b, (pc, sc_), st in zip(films[1:], counts[1:], stats[1:]):
            for p in b:
                assert (bits(b[p]) == bits(films[0][p])).all()
            assert (pc == counts[0][0]).all() and (sc_ == counts[0][1]).all()
            for key in ("rays_extension", "rays_shadow_env", "rays_shadow_light", "hit_bounces", "pixel_samples"):
                assert st[key] == stats[0][key], key
        assert stats[1]["kernel_launches"] < stats[0]["kernel_launches"]
    with pytest.raises(R.ElevenError):
        R.Renderer(rng_mode=R.RNG_REFERENCE, env_mode=R.ENV_CDF, flags=0, wave_spp=4)     # the reference stream is sequential per pixel
    with pytest.raises(R.ElevenError):
        R.Renderer(**dict(R.FAST, wave_spp=3))


def test_anyhit_light_shadows_equal_closest_hit_light_shadows(scenes):
    """ELEVEN_FLAG_ANYHIT_LIGHT_SHADOWS: point-light shadow rays stop at the first occluder in front of the light instead of
    taking the reference's closest hit (S/kernel.cu:193-197).  Same random numbers, so the film is identical except where
    an occluder sits within the shadow-terminator shift + 1 mm of the light itself."""
    sc = scenes["cornell"]
    base = dict(R.FAST); base["flags"] &= ~R._capi.FLAG_ANYHIT_LIGHT_SHADOWS
    a = R.Renderer(**base).render_setup(sc); a.render_cuda(16)
    b = R.Renderer(**R.FAST).render_setup(sc); b.render_cuda(16)
    fa, fb = a.film()[..., :3], b.film()[..., :3]
    assert (bits(fa) == bits(fb)).all(-1).mean() >= 0.999
    assert a.stats()["rays_shadow_light"] == b.stats()["rays_shadow_light"] > 0
    a.close(); b.close()


def test_packed_material_maps_do_not_change_the_film(scenes, monkeypatch):
    """The four 8-bit maps of a material are interleaved into one 8-byte record per texel at upload (DevPackedMaps): one
    gather per hit instead of four.  Same texel bytes through the same decode tables: the film must be bit-identical."""
    sc = scenes["clock"]
    a = R.Renderer(**R.PARITY).render_setup(sc); a.render_cuda(3)
    monkeypatch.setenv("ELEVEN_NO_PACKED_MAPS", "1")
    b = R.Renderer(**R.PARITY).render_setup(sc); b.render_cuda(3)
    monkeypatch.delenv("ELEVEN_NO_PACKED_MAPS")
    fa, fb = a.get_buffers()[0], b.get_buffers()[0]
    for p in fa:
        assert (bits(fa[p]) == bits(fb[p])).all()
    a.close(); b.close()


def test_env_alias_matches_cdf_distribution(scenes):
    sc = scenes["clock"]
    a = R.Renderer(rng_mode=R.RNG_FAST, env_mode=R.ENV_CDF, flags=0).render_setup(sc); a.render_cuda(96)
    b = R.Renderer(rng_mode=R.RNG_FAST, env_mode=R.ENV_ALIAS, flags=0, seed=9).render_setup(sc); b.render_cuda(96)
    ia, ib = a.film()[..., :3], b.film()[..., :3]
    assert abs(ia.mean() - ib.mean()) / ia.mean() < 0.03
    a.close(); b.close()


def test_resolve_rgba8_uses_reference_output_curve(scenes):
    sc = scenes["cornell"]
    r = R.Renderer(**R.PARITY).render_setup(sc); r.render_cuda(2)
    f = r.film()
    out = r.resolve_rgba8()
    x = np.clip(f, 0, 1).astype(np.float64)
    exp = np.array([O.lib().orc_fastpow(float(v), 1.0 / 2.2) * 255 for v in x.ravel()[:4096]]).astype(np.uint8)
    assert (out.ravel()[:4096] == exp).all()
    r.close()


def test_error_behaviour():
    r = R.Renderer(**R.PARITY)
    with pytest.raises(R.ElevenError):
        r.render_cuda(1)                       # ELEVEN_ERR_STATE: render before upload
    sc = S.cornell_box(16, env=(0, 0, 0), env_size=(8, 8))
    with pytest.raises(R.ElevenError):
        r.render_setup(sc)                     # black environment: the reference hangs (F10); we refuse
    sc = S.cornell_box(16, env_size=(8, 8))
    r.render_setup(sc)
    with pytest.raises(R.ElevenError):
        r.render_cuda(-1)
    with pytest.raises(R.ElevenError):
        R.Renderer(rng_mode=R.RNG_REFERENCE, flags=0, device=99)
    r.close()


def test_empty_and_degenerate_scenes():
    sc = S.cornell_box(16, env_size=(8, 8))
    sc.tris = sc.tris[:0]
    r = R.Renderer(**R.PARITY).render_setup(sc)
    hits = r.trace_closest(MG.ray_batch(S.cornell_box(16, env_size=(8, 8)), 64, 64, 0))
    assert (hits["tri"] == -1).all()
    r.render_cuda(2)
    assert np.allclose(r.film()[..., :3], 0.01, atol=1e-6)     # every path escapes into the constant environment
    assert r.trace_closest(np.zeros((0, 6), np.float32)).shape == (0,)
    r.close()
    # one degenerate (zero-area) triangle + one real one
    sc = S.cornell_box(16, env_size=(8, 8))
    sc.tris = sc.tris[:2].copy()
    sc.tris["vertices"][0] = sc.tris["vertices"][0][0]
    orc = O.Oracle(sc)
    rays = MG.ray_batch(S.cornell_box(16, env_size=(8, 8)), 256, 256, 0)
    r = R.Renderer(**R.PARITY).render_setup(sc)
    exact, tie, slab, bad = classify_hits(r.trace_closest(rays), orc.trace(rays, 0), orc.trace(rays, 1))
    assert bad.sum() == 0
    r.close(); orc.close()
