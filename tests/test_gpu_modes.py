"""GPU tests that pin the BENCHMARKED configuration flag by flag, the snapshot stream and the native film reduce.

The fast configuration is PARITY plus: counter RNG, alias-table environment sampling, dead-path termination, null-NEE skipping,
fast-math shading, any-hit light shadows.  Each one is isolated here:
  * null-NEE skipping and dead-path termination do not touch the estimator at all: with the same random numbers the BEAUTY film is
    BIT-IDENTICAL with and without them (asserted on the counter RNG, whose draws are keyed by (pixel, sample, bounce) and therefore
    do not shift when a path stops early; null-NEE skipping also under the reference XORWOW stream);
  * the alias table and fast-math shading change rounding / the sampling map, not the expectation: each alone, on top of PARITY,
    against the ORACLE;
  * any-hit light shadows: tests/test_gpu_parity.py::test_anyhit_light_shadows_equal_closest_hit_light_shadows.
"""
import threading
import time

import numpy as np
import pytest

import make_golden as MG
import oracle_lib as O
from gpu_metrics import record
from tfg_pathtracer_b200 import _capi, renderer as R
from tfg_pathtracer_b200 import scenes as S

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def scenes():
    d = MG.golden_scenes()
    d["zoo"] = S.material_zoo(lights=0)
    return d


BASE_FAST_RNG = dict(rng_mode=R.RNG_FAST, env_mode=R.ENV_CDF, hit_mode=R.HIT_KEY, flags=0)


@pytest.mark.parametrize("name", ["clock", "zoo", "cornell"])
def test_skip_null_nee_is_bit_identical(scenes, name):
    sc = scenes[name]
    for base in (BASE_FAST_RNG, R.PARITY):
        a = R.Renderer(**base).render_setup(sc); a.render_cuda(6)
        cfg = dict(base); cfg["flags"] = cfg["flags"] | _capi.FLAG_SKIP_NULL_NEE
        b = R.Renderer(**cfg).render_setup(sc); b.render_cuda(6)
        fa, pa = a.get_buffers(); fb, pb = b.get_buffers()
        for p in fa:
            assert (bits(fa[p]) == bits(fb[p])).all(), (name, p)
        assert (pa == pb).all() and (a.get_sample_counts() == b.get_sample_counts()).all()
        sa, sb = a.stats(), b.stats()
        assert sb["rays_extension"] == sa["rays_extension"]
        if len(sc.lights) == 0 and name != "zoo":                 # with lights or emission nothing is skippable
            assert sb["rays_shadow_env"] < sa["rays_shadow_env"]
        record("skip_null_nee_%s_rng%d" % (name, base["rng_mode"]), shadow_rays_without=sa["rays_shadow_env"], shadow_rays_with=sb["rays_shadow_env"])
        a.close(); b.close()


@pytest.mark.parametrize("name", ["clock", "zoo", "cornell", "grid"])
def test_dead_path_termination_is_bit_identical(scenes, name):
    """A path whose throughput is exactly 0 adds thr * (...) = 0 at every later bounce: stopping it leaves the radiance sum as it is —
    with ONE exception, and the test pins it: a later bounce of the dead path can produce 0 * inf = NaN (a zero pdf in a division),
    which makes the reference DROP the whole sample (S/kernel.cu:449: NaN samples are not accumulated and not counted).  The
    terminated path never gets there, so that sample is kept with the radiance it had.  Hence: every pixel is bit-identical in
    all passes unless its accepted-sample count differs, the terminated render never counts FEWER samples, and the samples involved
    are few (recorded; < 2 % of all samples on the scene that provokes it most, the displaced grid)."""
    sc = scenes[name]
    spp = 8
    a = R.Renderer(**BASE_FAST_RNG).render_setup(sc); a.render_cuda(spp)
    cfg = dict(BASE_FAST_RNG); cfg["flags"] = R.FLAG_TERMINATE_DEAD_PATHS
    b = R.Renderer(**cfg).render_setup(sc); b.render_cuda(spp)
    fa, fb = a.film(), b.film()
    same = (bits(fa) == bits(fb)).all(-1)
    ca, cb = a.get_sample_counts().reshape(same.shape), b.get_sample_counts().reshape(same.shape)
    m = record("terminate_" + name, identical_fraction=same.mean(), count_diff_pixels=int((ca != cb).sum()),
               nan_dropped_samples_without=int(spp * ca.size - ca.sum()), nan_dropped_samples_with=int(spp * cb.size - cb.sum()),
               sample_fraction_kept_instead_of_dropped=float((cb.astype(np.int64) - ca).sum() / (spp * ca.size)),
               ext_rays_without=a.stats()["rays_extension"], ext_rays_with=b.stats()["rays_extension"])
    assert (same | (ca != cb)).all(), m                             # equal counts => identical bits
    assert (cb >= ca).all(), m                                      # termination only ever KEEPS a sample the other render dropped
    assert (cb.astype(np.int64) - ca).sum() <= 0.02 * spp * ca.size, m
    for p in (R.PASS_NORMAL, R.PASS_TANGENT, R.PASS_BITANGENT):
        assert ((bits(a.film(p)) == bits(b.film(p))).all(-1) | (ca != cb)).all()
    assert b.stats()["rays_extension"] < a.stats()["rays_extension"]
    a.close(); b.close()


def _block_stats(img, ref):
    H, W = img.shape[:2]
    blk = lambda x: x[:H // 8 * 8, :W // 8 * 8].reshape(H // 8, 8, W // 8, 8, 3).mean((1, 3))
    rel = np.abs(blk(img) - blk(ref)) / (blk(ref) + 0.02)
    return float(np.median(rel)), float(np.percentile(rel, 95))


@pytest.mark.parametrize("name", ["clock", "zoo"])
def test_fast_math_alone_against_the_oracle(scenes, name):
    """PARITY + ELEVEN_FLAG_FAST_MATH (same XORWOW stream, same CDF search): MUFU approximations perturb every shading value by
    ~1e-6 relative, so a pixel either agrees to ~1e-4 or one of its paths took a different branch / texel / triangle and the pixel
    differs by Monte-Carlo noise; the mean must not move."""
    sc = scenes[name]
    spp = 8
    orc = O.Oracle(sc); orc.render(spp)
    ref = orc.film(0)[..., :3]
    cfg = dict(R.PARITY); cfg["flags"] = _capi.FLAG_FAST_MATH
    r = R.Renderer(**cfg).render_setup(sc); r.render_cuda(spp)
    img = r.film()[..., :3]
    close = (np.abs(img - ref) <= 1e-3 + 2e-3 * np.abs(ref)).all(-1)
    med, p95 = _block_stats(img, ref)
    m = record("fast_math_alone_" + name, close_fraction=close.mean(), mean_ours=img.mean(), mean_oracle=ref.mean(), median_block_rel=med, p95_block_rel=p95)
    assert close.mean() >= 0.90, m
    assert abs(img.mean() - ref.mean()) / ref.mean() < 0.005, m
    assert med < 0.01, m
    r.close(); orc.close()


@pytest.mark.parametrize("name", ["clock", "zoo"])
def test_alias_table_alone_against_the_oracle(scenes, name):
    """PARITY + ELEVEN_ENV_ALIAS: another map from uniforms to texels.  The alias table draws texel i with probability w_i, the pdf
    the estimator divides by (test_gpu_shading.py checks that distribution against the oracle's CDF).  The REFERENCE's binarySearch
    does not: for about half of the uniforms it returns the texel after the one whose CDF interval holds r (S/HDRI.hpp:130-142,
    measured on the oracle: 50.4 % of the picks), while HDRI::pdf is evaluated for the returned texel — a biased estimator wherever
    neighbouring texels differ (1.2 % of the image mean on the 128x64 environment of the zoo scene, nothing measurable on smooth ones).
    So the yardstick for the alias mode is the oracle with EXACT CDF inversion (oracle_lib.Oracle.set_env_search(True): every other
    line of the estimator is the reference's); the reference-search oracle is rendered too and the gap between the two recorded."""
    sc = scenes[name]
    spp = 64
    orc = O.Oracle(sc); orc.render(spp)
    ref_search = orc.film(0)[..., :3].copy()
    orc.reset(); orc.set_env_search(True); orc.render(spp)
    ref = orc.film(0)[..., :3]
    cfg = dict(R.PARITY); cfg["env_mode"] = R.ENV_ALIAS
    r = R.Renderer(**cfg).render_setup(sc); r.render_cuda(spp)
    img = r.film()[..., :3]
    med, p95 = _block_stats(img, ref)
    # noise yardstick: image means of independent realisations of the SAME estimator (counter RNG, four seeds, alias table)
    means = []
    for seed in range(4):
        f = R.Renderer(rng_mode=R.RNG_FAST, env_mode=R.ENV_ALIAS, hit_mode=R.HIT_KEY, flags=0, seed=100 + seed).render_setup(sc)
        f.render_cuda(spp); means.append(float(f.film()[..., :3].mean())); f.close()
    sigma = float(np.std(means, ddof=1))
    m = record("alias_alone_" + name, mean_alias=img.mean(), mean_oracle_exact_inversion=ref.mean(), mean_oracle_reference_search=ref_search.mean(),
               reference_search_bias=float((ref_search.mean() - ref.mean()) / ref.mean()), sigma_of_image_mean=sigma, seed_means=means,
               median_block_rel=med, p95_block_rel=p95)
    assert abs(img.mean() - ref.mean()) < max(5.0 * np.sqrt(2.0) * sigma, 0.002 * ref.mean()), m
    assert abs(np.mean(means) - ref.mean()) < max(5.0 * sigma, 0.002 * ref.mean()), m
    assert med < 0.05, m
    r.close(); orc.close()


def test_snapshot_stream_serves_the_film_while_a_render_is_running(scenes):
    """S/main.cpp:132-184 polls getBuffers / getSamples every 100 ms on a second stream while renderCuda runs in a worker thread
    (S/kernel.cu:688-710).  Same contract: a thread sits in ONE eleven_render call; this thread polls the film and the sample count
    on the context's snapshot stream and must get answers BEFORE the render returns, non-decreasing, each a valid partial film."""
    sc = S.clock_standin(tex_res=64, xres=960, yres=540, env_size=(256, 128))
    total = 512
    r = R.Renderer(**R.FAST).render_setup(sc)
    r.render_cuda(16); r.reset()                                   # warm-up (module load, first launches)
    done = {}

    def work():
        t0 = time.perf_counter(); r.render_cuda(total); done["t"] = time.perf_counter() - t0

    th = threading.Thread(target=work)
    polls = []
    t_start = time.perf_counter()
    th.start()
    while th.is_alive():
        t0 = time.perf_counter()
        n = r.get_samples()
        f = r.film()
        polls.append((time.perf_counter() - t_start, time.perf_counter() - t0, n, float(f[..., :3].mean()), bool(np.isfinite(f).all()), th.is_alive()))
        time.sleep(0.002)
    th.join()
    final = r.film()
    during = [p for p in polls if p[5]]                            # polls that returned while the render call was still running
    partial = [p for p in during if 0 < p[2] < total]
    m = record("snapshot_polling", render_s=done["t"], polls=len(polls), polls_during_render=len(during), partial_films=len(partial),
               max_poll_s=max(p[1] for p in polls), samples_seen=[p[2] for p in polls][:40])
    assert len(partial) >= 2, m                                   # films of a render in progress were served
    ns = [p[2] for p in polls]
    assert all(b >= a for a, b in zip(ns, ns[1:])), "sample counts must not go backwards"
    assert all(p[4] for p in polls)
    assert max(p[1] for p in during) < 0.5 * done["t"], m        # a poll does not wait for the render to end
    # a partial film is a converging picture of the same scene: its mean is close to the final one
    assert all(abs(p[3] - final[..., :3].mean()) / final[..., :3].mean() < 0.1 for p in partial)
    assert r.get_samples() == total
    # and the finished film is exactly what an unpolled render produces
    q = R.Renderer(**R.FAST).render_setup(sc); q.render_cuda(total)
    assert (bits(q.film()) == bits(final)).all()
    q.close(); r.close()


def test_native_film_reduce_on_one_gpu(scenes):
    """eleven_reduce_film without peers (no communicator, and a 1-rank NCCL communicator): the reduced film is the local film, the
    local film is untouched, and repeating the step while rendering goes on never double-counts (bench.py:255)."""
    sc = scenes["cornell"]
    for with_comm in (False, True):
        r = R.Renderer(**R.FAST).render_setup(sc)
        if with_comm:
            r.comm_init_rank(r.comm_unique_id(), 1, 0)
        with pytest.raises(R.ElevenError):
            r.film_reduced()                                       # nothing reduced yet
        for step in range(3):
            r.render_cuda(5)
            before = r.film()
            r.reduce_film(0, all_passes=(step == 2))
            assert (bits(r.film()) == bits(before)).all()
            assert (bits(r.film_reduced()) == bits(before)).all()
            assert (r.get_sample_counts_reduced() == r.get_sample_counts()).all() and r.get_sample_counts().max() == 5 * (step + 1)
            assert (r.resolve_rgba8_reduced() == r.resolve_rgba8()).all()
            if step == 2:
                for p in (R.PASS_NORMAL, R.PASS_TANGENT, R.PASS_BITANGENT):
                    assert (bits(r.film_reduced(p)) == bits(r.film(p))).all()
            else:
                with pytest.raises(R.ElevenError):
                    r.film_reduced(R.PASS_NORMAL)                  # all_passes = 0 reduced BEAUTY only
        st = r.stats()
        assert st["reduce_calls"] == 3 and st["reduce_ms"] > 0
        with pytest.raises(R.ElevenError):
            r.reduce_film(1)                                       # root out of range
        r.close()


def test_native_film_reduce_across_two_gpus(scenes):
    """Two contexts on two devices of this box, one NCCL communicator (eleven_comm_init_all), samples split even / odd: the reduced
    film on device 0 equals the film of one context rendering all samples, up to float summation order."""
    import ctypes as C
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = scenes["clock"]
    parts = []
    for g in range(2):
        cfg = dict(R.FAST); cfg.update(device=g, sample_offset=g, sample_stride=2)
        parts.append(R.Renderer(**cfg).render_setup(sc))
    arr = (C.c_void_p * 2)(parts[0].h, parts[1].h)
    parts[0]._ck(parts[0].L.eleven_comm_init_all(arr, 2))
    for p in parts:
        p.render_cuda(8)
    th = [threading.Thread(target=p.reduce_film, args=(0, True)) for p in parts]
    [t.start() for t in th]; [t.join() for t in th]
    one = R.Renderer(**R.FAST).render_setup(sc); one.render_cuda(16)
    assert (parts[0].get_sample_counts_reduced() == one.get_sample_counts()).all()
    for p in (R.PASS_BEAUTY, R.PASS_NORMAL):
        np.testing.assert_allclose(parts[0].film_reduced(p), one.film(p), rtol=2e-5, atol=1e-6)
    with pytest.raises(R.ElevenError):
        parts[1].film_reduced()                                    # only the root holds the reduced film
    for p in parts + [one]:
        p.close()
