#!/usr/bin/env python
"""Convergence check of BASELINE.json's north_star, part 2: "at high spp the image must converge to the same RMSE against a
converged reference".  Runs on the GPU box.

  converged reference  = the UNMODIFIED reference CUDA build (oracle/_ref/eleven_ref_headless_precise) at --ref-spp
  RMSE_ref(N)          = reference at N spp vs converged reference
  RMSE_ours(N)         = our FAST mode (counter RNG, alias-table env sampling, dead-path termination, null-NEE skipping,
                         fast-math shading) at N spp vs the same converged reference
  bias                 = |mean(ours@Nmax) - mean(converged)| / mean(converged)

Scene: the ClockCC0 stand-in at a reduced resolution (the reference renders ~1 M pixel-samples/s at small frames), taken
from the reference's own loader dump so that both renderers see identical inputs (MikkTSpace tangents, stb texel decode).
Prints one JSON line and writes gpurun_out/convergence.json.
"""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from tfg_pathtracer_b200 import renderer as R, scenes as S  # noqa: E402


def run_ref(binp, sdir, spp, prefix, dump=None):
    cmd = [binp, sdir, str(spp), prefix] + (["--dump-scene", dump] if dump else [])
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=sdir)
    if p.returncode != 0:
        raise SystemExit("reference failed: " + (p.stderr or p.stdout)[-500:])
    info = json.loads(open(prefix + ".json").read())
    img = np.fromfile(prefix + ".beauty.f32", np.float32).reshape(info["height"], info["width"], 4)[..., :3]
    return img, info


def rmse(a, b):
    return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--tex", type=int, default=256)
    ap.add_argument("--ref-spp", type=int, default=2048)
    ap.add_argument("--levels", default="16,64,256")
    ap.add_argument("--scene", default="clock", choices=["clock", "cornell_nolight"])
    a = ap.parse_args()
    levels = [int(x) for x in a.levels.split(",")]
    work = "/tmp/convergence_%d" % os.getpid()
    if a.scene == "clock":
        sc = S.clock_standin(tex_res=a.tex, xres=a.res, yres=a.res * 9 // 16, env_size=(1024, 512))
        env_color = None
    else:
        sc = S.cornell_box(a.res, light=False, tilt=(3.0, 7.0, 2.0), box_gap=0.002)
        env_color = (0.01, 0.01, 0.01)
    S.write_reference_scene_dir(sc, work, env_color=env_color)
    binp = os.path.join(ROOT, "oracle", "_ref", "eleven_ref_headless_precise")
    dump = os.path.join(work, "scene.flat")
    conv, info = run_ref(binp, work, a.ref_spp, os.path.join(work, "conv"), dump)
    sd = S.load_flat(dump)
    out = {"scene": a.scene, "width": info["width"], "height": info["height"], "converged_spp": a.ref_spp,
           "converged_mean": float(conv.mean()), "levels": {}}
    for n in levels:
        ref_n, _ = run_ref(binp, work, n, os.path.join(work, "ref%d" % n))
        f = R.Renderer(**R.FAST).render_setup(sd)
        f.render_cuda(n)
        ours = f.film()[..., :3]
        f.close()
        p = R.Renderer(**R.PARITY).render_setup(sd)
        p.render_cuda(n)
        par = p.film()[..., :3]
        p.close()
        e_ref, e_ours, e_par = rmse(ref_n, conv), rmse(ours, conv), rmse(par, conv)
        # The reference at N spp is a PREFIX of its own converged render (same per-pixel streams), so its error against it is
        # sqrt(var (1/N - 1/M)), while an independent stream gives sqrt(var (1/N + 1/M)): compare per-sample variances.
        M_ = float(a.ref_spp)
        var_ref = e_ref ** 2 / (1.0 / n - 1.0 / M_)
        var_fast = e_ours ** 2 / (1.0 / n + 1.0 / M_)
        out["levels"][str(n)] = {"rmse_reference": e_ref, "rmse_ours_fast": e_ours, "rmse_ours_parity": e_par,
                                 "ratio_fast": e_ours / e_ref, "ratio_parity": e_par / e_ref,
                                 "expected_ratio_for_an_independent_stream": float(np.sqrt((1.0 / n + 1.0 / M_) / (1.0 / n - 1.0 / M_))),
                                 "per_sample_variance_reference": var_ref, "per_sample_variance_ours_fast": var_fast,
                                 "variance_ratio_fast_over_reference": var_fast / var_ref,
                                 "mean_reference": float(ref_n.mean()), "mean_ours_fast": float(ours.mean())}
    hi = R.Renderer(**R.FAST).render_setup(sd)
    hi.render_cuda(a.ref_spp)
    ours_hi = hi.film()[..., :3]
    hi.close()
    out["bias_fast_vs_converged"] = abs(float(ours_hi.mean()) - float(conv.mean())) / float(conv.mean())
    out["rmse_fast_converged_vs_reference_converged"] = rmse(ours_hi, conv)
    blk = lambda x: x[: x.shape[0] // 8 * 8, : x.shape[1] // 8 * 8].reshape(x.shape[0] // 8, 8, x.shape[1] // 8, 8, 3).mean((1, 3))
    out["max_block8_rel_diff_converged"] = float((np.abs(blk(ours_hi) - blk(conv)) / (blk(conv) + 1e-3)).max())
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "convergence_%s.json" % a.scene), "w"), indent=1)


if __name__ == "__main__":
    main()
