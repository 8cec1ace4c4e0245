#!/usr/bin/env python
"""Runs on the GPU box: renders the same scene with (a) the UNMODIFIED reference CUDA renderer
(oracle/_ref/eleven_ref_headless_{precise,fast}, built from /root/reference by oracle/Makefile), (b) the CPU oracle and
(c) our CUDA path, from the scene exactly as the reference's loader produced it (its flat dump: MikkTSpace tangents,
stb decode), and reports agreement + timings.  With --save, writes the reference image + the flat scene as a golden
fixture candidate (tests/golden/ref_cuda_<name>.npz)."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from tfg_pathtracer_b200 import renderer as R, scenes as S  # noqa: E402


def build_scene(name, res, tex):
    if name == "cornell":
        return S.cornell_box(res), (0.01, 0.01, 0.01)
    if name == "cornell_nolight":
        return S.cornell_box(res, light=False), (0.01, 0.01, 0.01)
    if name == "clock":
        return S.clock_standin(tex_res=tex, xres=res, yres=res * 9 // 16, env_size=(max(64, tex * 2), max(32, tex))), None
    if name == "clock_full":
        return S.clock_standin(tex_res=tex), None
    if name == "grid":
        return S.displaced_grid(res, xres=480, yres=270, env_size=(512, 256)), None
    raise SystemExit("unknown scene " + name)


def within(a, b, atol=1e-3, rtol=1e-3):
    return float((np.abs(a - b) <= atol + rtol * np.abs(b)).all(-1).mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell")
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--tex", type=int, default=64)
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--flavour", default="precise")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--oracle", action="store_true", help="also render with the CPU oracle (slow at high res)")
    ap.add_argument("--save", action="store_true")
    ap.add_argument("--ours-spp", type=int, default=0, help="also time our fast mode at this spp")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    work = os.path.join("/tmp", "refcmp_%s_%d" % (a.scene, os.getpid()))
    sc, env_color = build_scene(a.scene, a.res, a.tex)
    t0 = time.time()
    S.write_reference_scene_dir(sc, work, env_color=env_color)
    t_write = time.time() - t0
    binp = os.path.join(ROOT, "oracle", "_ref", "eleven_ref_headless_" + a.flavour)
    prefix = os.path.join(work, "ref")
    dump = os.path.join(work, "scene.flat")
    ext = ["--external-textures"] if a.scene == "clock_full" else []
    t0 = time.time()
    p = subprocess.run([binp, work, str(a.spp), prefix, "--dump-scene", dump] + ext, capture_output=True, text=True, cwd=work)
    t_ref = time.time() - t0
    if p.returncode != 0:
        print(p.stdout[-2000:], p.stderr[-2000:])
        raise SystemExit("reference run failed rc=%d" % p.returncode)
    info = json.loads(open(prefix + ".json").read())
    W, H = info["width"], info["height"]
    ref = {k: np.fromfile(prefix + "." + k + ".f32", np.float32).reshape(H, W, 4) for k in ("beauty", "normal", "tangent", "bitangent")}
    ref_pc = np.fromfile(prefix + ".pathcount.i32", np.int32)
    sd = S.load_flat(dump)
    report = {"scene": a.scene, "flavour": a.flavour, "spp": a.spp, "reference": info, "write_scene_s": t_write, "ref_wall_s": t_ref,
              "tris": int(len(sd.tris))}

    r = R.Renderer(**R.PARITY).render_setup(sd)
    r.render_cuda(a.spp)
    bufs, pc = r.get_buffers()
    st = r.stats()
    report["ours_parity"] = {"render_ms": st["render_ms"], "samples_per_s": W * H * a.spp / (st["render_ms"] * 1e-3),
                             "bvh_build_ms": st["bvh_build_ms"], "bvh_nodes": st["bvh_nodes"], "key_slack": st["key_slack"],
                             "rays": int(st["rays_extension"] + st["rays_shadow_env"] + st["rays_shadow_light"]),
                             "beauty_within_tol_vs_ref": within(bufs[0][..., :3], ref["beauty"][..., :3]),
                             "normal_within_1e-5_vs_ref": within(bufs[2][..., :3], ref["normal"][..., :3], 1e-5, 1e-5),
                             "tangent_within_1e-5_vs_ref": within(bufs[3][..., :3], ref["tangent"][..., :3], 1e-5, 1e-5),
                             "pathcount_equal_vs_ref": float((pc == ref_pc).mean()),
                             "mean_ours": float(bufs[0][..., :3].mean()), "mean_ref": float(ref["beauty"][..., :3].mean())}
    r.close()
    if a.oracle:
        import oracle_lib as O
        orc = O.Oracle(sd)
        t0 = time.time(); orc.render(a.spp, threads=os.cpu_count() or 8); t_or = time.time() - t0
        of = orc.film(0)
        report["oracle"] = {"render_s": t_or, "threads": os.cpu_count(), "beauty_within_tol_vs_ref": within(of[..., :3], ref["beauty"][..., :3]),
                            "normal_within_1e-5_vs_ref": within(orc.film(2)[..., :3], ref["normal"][..., :3], 1e-5, 1e-5),
                            "ours_within_tol_vs_oracle": within(bufs[0][..., :3], of[..., :3]),
                            "pathcount_equal_vs_ref": float((orc.counts()[1].astype(np.int32) == ref_pc).mean())}
        orc.close()
    if a.ours_spp:
        f = R.Renderer(**R.FAST).render_setup(sd)
        f.render_cuda(2); f.reset()
        f.render_cuda(a.ours_spp)
        st = f.stats()
        report["ours_fast"] = {"spp": a.ours_spp, "render_ms": st["render_ms"], "samples_per_s": W * H * a.ours_spp / (st["render_ms"] * 1e-3),
                               "mrays_per_s": (st["rays_extension"] + st["rays_shadow_env"] + st["rays_shadow_light"]) / (st["render_ms"] * 1e3),
                               "mean": float(f.film()[..., :3].mean())}
        f.close()
    if a.save:
        np.savez_compressed(os.path.join(a.out, "ref_cuda_%s.npz" % a.scene), beauty=ref["beauty"], normal=ref["normal"], tangent=ref["tangent"],
                            bitangent=ref["bitangent"], pathcount=ref_pc, spp=a.spp, flavour=a.flavour,
                            scene_flat=np.frombuffer(open(dump, "rb").read(), np.uint8))
    print(json.dumps(report))
    with open(os.path.join(a.out, "refcmp_%s_%s.json" % (a.scene, a.flavour)), "w") as fjs:
        json.dump(report, fjs, indent=1)


if __name__ == "__main__":
    main()
