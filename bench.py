#!/usr/bin/env python
"""bench.py — samples/s (and Mrays/s) of the path-tracing hot path on the ClockCC0 stand-in, 1..8 B200s.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the ClockCC0 STAND-IN scene
(125 281 triangles, 12 x 4096^2 8-bit maps, 4096x2048 HDRI, depth of field; the real asset is not in the reference
checkout) at 1920x1080.  One step = `--spp-per-step` samples of every pixel on every GPU (weak scaling:
GPU g renders the global samples g, g+N, g+2N, ...), followed for N > 1 by one NCCL reduce of the per-GPU film sums +
sample counts to rank 0 over NVLink.  value = pixel-samples rendered by all ranks / time.

  value     fast path (counter RNG, alias-table env sampling, dead-path termination), scene resident in HBM
  e2e       same metric through the C ABI with host buffers: per step a camera upload (H2D), the render, and the
            BEAUTY film read back to host memory (D2H), all inside the timed region
  roofline  the closest-hit kernel: algorithmic bytes/ray (80 B x nodes + 48 B x triangles + 48 B ray in / hit out,
            counted by a counters build of the same kernel on the same rays) / its CUDA-event time, vs measured HBM copy
  cpu_baseline   the CPU oracle (port of the reference algorithm) on the host cores, bounded sample

`--impl reference` times the UNMODIFIED reference renderer.  The reference has no CPU render path, so —
as BASELINE.json's north_star prescribes — this arm runs the reference CUDA build (oracle/_ref/eleven_ref_headless_fast,
the flavour its author shipped: -use_fast_math) on ONE B200, labelled as such.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

CACHE = os.environ.get("ELEVEN_BENCH_CACHE", "/tmp/eleven_bench_cache")


# ------------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def get_scene(args, need_dir):
    """ClockCC0 stand-in, generated once per box and cached (flat container for us/oracle, scene dir for the reference)."""
    from tfg_pathtracer_b200 import scenes as S
    workload = getattr(args, "workload", "clock")
    tag = "clock_t%d_%dx%d" % (args.tex, args.width, args.height) if workload == "clock" else "grid_n%d_%dx%d" % (args.grid, args.width, args.height)
    flat = os.path.join(CACHE, tag + ".flat")
    sdir = os.path.join(CACHE, tag + "_dir")
    os.makedirs(CACHE, exist_ok=True)
    lock = os.path.join(CACHE, tag + ".lock")
    # one generator per box; other ranks wait
    while True:
        try:
            fd = os.open(lock, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
            os.close(fd)
            break
        except FileExistsError:
            if os.path.exists(flat + ".done") and (not need_dir or os.path.exists(sdir + ".done")):
                return flat, sdir
            time.sleep(0.5)
    try:
        sc = None
        if not os.path.exists(flat + ".done"):
            sc = S.clock_standin(tex_res=args.tex, xres=args.width, yres=args.height) if workload == "clock" else \
                S.displaced_grid(args.grid, xres=args.width, yres=args.height)
            S.save_flat(sc, flat)
            open(flat + ".done", "w").close()
        if need_dir and not os.path.exists(sdir + ".done"):
            sc = sc or S.load_flat(flat)
            sc.object_names, sc.material_names = ["clock", "table", "plant"], ["clock_mat", "table_mat", "plant_mat"]
            S.write_reference_scene_dir(sc, sdir)
            open(sdir + ".done", "w").close()
    finally:
        os.remove(lock)
    return flat, sdir


# ------------------------------------------------------------------------------------------------------
def run_reference(args, rank, out):
    if rank != 0:
        return
    binp = os.path.join(ROOT, "oracle", "_ref", "eleven_ref_headless_fast")
    base = {"impl": "reference", "metric": "samples_per_second", "unit": "pixel-samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if not os.path.exists(binp):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/eleven_ref_headless_fast not built (needs /root/reference at build time)"}), file=out, flush=True)
        return
    flat, sdir = get_scene(args, need_dir=True)
    spp_w, spp_t = args.warmup * args.ref_spp_per_step, args.steps * args.ref_spp_per_step
    sm = ClockSampler(0)
    sm.start()
    t0 = time.time()
    p = subprocess.run([binp, sdir, str(spp_t), os.path.join(CACHE, "ref_out"), "--warmup", str(spp_w)], capture_output=True, text=True, cwd=sdir)
    wall = time.time() - t0
    clocks = sm.stop()
    if p.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable": "reference run failed rc=%d: %s" % (p.returncode, (p.stderr or p.stdout)[-300:].replace("\n", " "))}), file=out, flush=True)
        return
    info = json.loads(open(os.path.join(CACHE, "ref_out.json")).read())
    v = info["samples_per_s"]
    base.update({"value": v, "ms_per_step": info["render_ms"] / args.steps,
                 "config": {"workload": "ClockCC0 stand-in %dx%d, %d spp/step, reference renderer (CUDA build, -use_fast_math) on 1 B200" % (args.width, args.height, args.ref_spp_per_step),
                            "spp_per_step": args.ref_spp_per_step, "tris": info["tris"], "l2": "working set (2.4 GB float textures) larger than L2"},
                 "cpu_baseline": {"value": v, "unit": "pixel-samples/s", "cores": 0, "kind": "reference",
                                  "sample": "%d spp of the full frame on ONE B200 (the reference has no CPU render path, SURVEY F3; north_star: report its CUDA build, labelled)" % spp_t,
                                  "device": "1x B200, reference CUDA build"},
                 "e2e": {"value": v, "unit": "pixel-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "reference": {"load_ms": info["load_ms"], "setup_ms": info["setup_ms"], "render_ms": info["render_ms"], "kpaths_per_s": info["kpaths_per_s"],
                               "hit_bounces": info["hit_bounces"], "wall_s": wall},
                 "clocks": clocks, "gpu_launches": spp_t + spp_w})
    print(json.dumps(base), file=out, flush=True)


# ------------------------------------------------------------------------------------------------------
NCU_METRICS = ("smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
               "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum")


def ncu_counters(args, git_sha):
    """Hardware counters of ONE wave of this very workload, taken in THIS run (untimed, before the timed region): an ncu
    subprocess replays the traversal/shade kernels of one `--spp-per-step` wave (tools/profile_run.py loads the same cached
    scene, same mode, same seed: the counter-based RNG makes it the same rays as a bench step).  Returns per-kernel sums
    {kernel: {launches, warp_inst, thread_inst, dram_bytes, l2_bytes}} or None (+ reason) when ncu is unavailable."""
    import csv
    import shutil
    import tempfile
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    log = tempfile.mktemp(prefix="eleven_ncu_", suffix=".csv")
    cmd = [ncu, "--metrics", ",".join(NCU_METRICS), "--clock-control", "none", "--csv", "--log-file", log,
           "-k", "regex:k_extend|k_shade|k_shadowEnv|k_shadowLight|k_classify",
           sys.executable, os.path.join(ROOT, "tools", "profile_run.py"), "--spp", str(args.spp_per_step), "--tex", str(args.tex),
           "--width", str(args.width), "--height", str(args.height), "--mode", args.mode, "--hit-mode", args.hit_mode,
           "--workload", args.workload, "--grid", str(args.grid), "--wave-spp", str(args.wave_spp)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    except Exception as e:       # noqa: BLE001
        return None, "ncu failed to run: %r" % (e,)
    if p.returncode != 0 or not os.path.exists(log):
        return None, "ncu rc=%d: %s" % (p.returncode, (p.stderr or p.stdout)[-300:].replace("\n", " "))
    agg = {}
    try:
        rows = list(csv.reader(open(log)))
        hdr = next(r for r in rows if "Kernel Name" in r and "Metric Name" in r)
        ik, im, iv, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
        seen = {}
        for r in rows[rows.index(hdr) + 1:]:
            if len(r) <= max(ik, im, iv):
                continue
            name = r[ik].split("(")[0].replace("eleven::", "").replace("void ", "").split("<")[0].strip()
            a = agg.setdefault(name, {"launches": 0, "warp_inst": 0.0, "thread_inst": 0.0, "dram_bytes": 0.0, "l2_bytes": 0.0})
            if (name, r[ii]) not in seen:
                seen[(name, r[ii])] = 1
                a["launches"] += 1
            v = float(r[iv].replace(",", ""))
            m = r[im]
            if m == "smsp__inst_executed.sum":
                a["warp_inst"] += v
            elif m == "smsp__thread_inst_executed.sum":
                a["thread_inst"] += v
            elif m.startswith("dram__bytes"):
                a["dram_bytes"] += v
            elif m.startswith("lts__t_sectors"):
                a["l2_bytes"] += 32.0 * v
    except Exception as e:       # noqa: BLE001
        return None, "cannot parse the ncu log: %r" % (e,)
    finally:
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            shutil.copy(log, os.path.join(ROOT, "gpurun_out", "bench_ncu_counters_%s.csv" % args.workload))
            os.remove(log)
        except Exception:        # noqa: BLE001
            pass
    if "k_extend" not in agg:
        return None, "no k_extend launch in the ncu log"
    return agg, "ncu subprocess of this run (git %s): %s over one %d-spp wave" % (git_sha, ", ".join(NCU_METRICS[:2]), args.spp_per_step)


def git_head():
    try:
        sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True, timeout=10).stdout.strip()
        if sha:
            return sha
    except Exception:            # noqa: BLE001
        pass
    try:                         # a gpurun snapshot has no .git: __graft_entry__.build() left the commit next to the library
        return json.load(open(os.path.join(ROOT, "tfg-pathtracer_b200", "_build_info.json")))["git"]
    except Exception:            # noqa: BLE001
        return "unknown"


def cpu_baseline(args, flat):
    import oracle_lib as O
    from tfg_pathtracer_b200 import scenes as S
    sc = S.load_flat(flat)
    threads = os.cpu_count() or 1
    t0 = time.time()
    orc = O.Oracle(sc)
    setup = time.time() - t0
    rows = max(8, min(sc.height, int(args.cpu_rows)))
    y0 = (sc.height - rows) // 2
    spp = max(1, int(args.cpu_spp))
    t0 = time.time()
    orc.render(spp, threads=threads, rows=(y0, y0 + rows))
    dt = time.time() - t0
    n = rows * sc.width * spp
    rc = orc.ray_counts()
    orc.close()
    return {"value": n / dt, "unit": "pixel-samples/s", "cores": threads, "kind": "port",
            "sample": "%d spp of %d central rows (%d pixel-samples) of the same frame, oracle/eleven_oracle.cpp, %.1f s (+%.1f s reference-style BVH build)" % (spp, rows, n, dt, setup),
            "mrays_per_s": float(rc.sum()) / dt / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=16)
    ap.add_argument("--ref-spp-per-step", type=int, default=4)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--tex", type=int, default=4096)
    ap.add_argument("--cpu-rows", type=int, default=1080)
    ap.add_argument("--cpu-spp", type=int, default=8)
    ap.add_argument("--wave-spp", type=int, default=0, help="samples of every pixel in flight per wave (0 = auto)")
    ap.add_argument("--bvh", default="", choices=["", "device", "host"], help="BVH8 builder (default: the mode's own: device for fast, host for parity)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="fast", choices=["fast", "parity"])
    ap.add_argument("--hit-mode", default="key", choices=["key", "min_t"], help="closest-hit ordering: the reference key (default) or classic min t")
    ap.add_argument("--workload", default="clock", choices=["clock", "grid"],
                    help="clock: ClockCC0 stand-in (BASELINE configs[2], the metric's configuration); grid: configs[3], the ~10 M-triangle displaced grid "
                         "(BVH8 + triangles ~560 MB >> L2: the workload whose traversal roof is HBM)")
    ap.add_argument("--grid", type=int, default=2237, help="grid workload: n x n vertices -> 2(n-1)^2 triangles (2237 -> 9 999 392)")
    ap.add_argument("--no-ncu", action="store_true", help="skip the live ncu counter pass (roofline.issue / traffic fall back to profiles/traffic.json, flagged)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON: anything native libraries print to fd 1 (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        run_reference(args, rank, json_out)
        return

    import torch
    import torch.distributed as dist
    from tfg_pathtracer_b200 import dist as D, renderer as R, scenes as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    flat, _ = get_scene(args, need_dir=False)
    sc = S.load_flat(flat)
    W, H = sc.width, sc.height
    mode = dict(R.FAST if args.mode == "fast" else R.PARITY)
    if args.hit_mode == "min_t":
        mode["hit_mode"] = R.HIT_MIN_T
    t0 = time.time()
    off, stride, _ = D.sample_plan(args.spp_per_step * world, rank, world)
    if args.mode == "fast":
        mode["wave_spp"] = args.wave_spp
    if args.bvh:
        mode["bvh_builder"] = R.BVH_DEVICE if args.bvh == "device" else R.BVH_HOST
    r = R.Renderer(device=local, sample_offset=off, sample_stride=stride, **mode).render_setup(sc)
    setup_s = time.time() - t0
    if world > 1:
        D.init_comm(r, rank, world)             # native NCCL communicator of the context; torch.distributed only carries the 128-byte id
    S_ = args.spp_per_step

    def step():
        r.render_cuda(S_)                       # blocking; its device time is measured by CUDA events on the context's stream (ElevenStats.render_ms)
        if world > 1:                           # the one exchange step: ONE ncclReduce of the film records (sum.xyz, count) to rank 0, enqueued by the
            r.reduce_film(0)                    # library on the same stream, into a buffer separate from the local film; timed by its own CUDA events

    def device_ms(st0):
        """Device time of the steps since the stats snapshot `st0`: render + reduce, CUDA events on the context's stream."""
        torch.cuda.synchronize()
        st = r.stats()
        return st["render_ms"] - st0["render_ms"] + st["reduce_ms"] - st0["reduce_ms"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- per-ray algorithmic bytes from a counters build of the same kernels on the same rays (untimed) ----
    rc_cfg = dict(mode)
    rc_cfg["flags"] = rc_cfg["flags"] | R.FLAG_COUNTERS
    sha = git_head()
    ncu, ncu_src = None, "skipped (--no-ncu)"
    if rank == 0:
        rc = R.Renderer(device=local, **rc_cfg).render_setup(sc)
        rc.render_cuda(S_)
        cst = rc.stats()
        rc.close()
        # ---- hardware counters of one wave of this workload, measured in this run (ncu subprocess; untimed) ----
        if not args.no_ncu:
            ncu, ncu_src = ncu_counters(args, sha)
            if ncu is None:
                print("bench.py: live ncu counters unavailable: %s" % ncu_src, file=sys.stderr)
    else:
        cst = None

    for _ in range(args.warmup):
        r.reset()
        step()
    barrier()
    # ---- timed: value -------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r.reset()
    st0 = r.stats()
    launches0 = st0["kernel_launches"]
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    dt_host = time.perf_counter() - t0
    dt = device_ms(st0) * 1e-3                                           # CUDA events; max over ranks below
    st = r.stats()
    # ---- timed: stage breakdown for the roofline (separate pass so that the events do not perturb `value`) ------
    tk_cfg = dict(mode)
    tk_cfg["flags"] = tk_cfg["flags"] | R.FLAG_TIME_KERNELS
    stage = None
    if rank == 0:
        rt = R.Renderer(device=local, **tk_cfg).render_setup(sc)
        rt.render_cuda(2)
        rt.reset()
        rt.render_cuda(S_)
        stage = rt.stats()
        rt.close()
    # ---- timed: e2e ---------------------------------------------------------------------------------------
    host_film = r.pinned_array((H, W, 4), np.float32)                 # page-locked host buffer for the per-step film read-back
    r.reset()
    barrier()
    t1 = time.perf_counter()
    for k in range(args.steps):
        r.set_camera(sc.camera)                                     # H2D: 56 bytes
        step()
        if rank == 0:                                               # resolve + D2H of the result: the film of all ranks (reduced) for N > 1
            if world > 1:
                r.film_reduced(R.PASS_BEAUTY, out=host_film)
            else:
                r._ck(r.L.eleven_get_film(r.h, R.PASS_BEAUTY, host_film.ctypes.data, W * H))
    barrier()
    dt_e2e = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([dt, dt_e2e, dt_host], device="cuda:%d" % local, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e, dt_host = float(t[0]), float(t[1]), float(t[2])
        rays = torch.tensor([st["rays_extension"], st["rays_shadow_env"], st["rays_shadow_light"]], device="cuda:%d" % local, dtype=torch.float64)
        dist.all_reduce(rays)
        rays_total = float(rays.sum())
    else:
        rays_total = float(st["rays_extension"] + st["rays_shadow_env"] + st["rays_shadow_light"])

    if rank == 0:
        total = float(W) * H * S_ * args.steps * world
        value = total / dt
        peak, peak_src = load_peaks()
        # algorithmic bytes of the dominant kernel: per extension ray 80 B x nodes + 48 B x triangles + 32 B ray in + 16 B hit out,
        # nodes / triangles counted for k_extend alone by the counters build on the same wave
        ext_rays_c = max(1, cst["rays_extension"])
        nodes_per_ray = cst["nodes_visited_extend"] / ext_rays_c
        tris_per_ray = cst["tris_tested_extend"] / ext_rays_c
        keys_per_ray = cst["key_evals"] / max(1, cst["rays_extension"] + cst["rays_shadow_env"] + cst["rays_shadow_light"])
        sh_rays_c = max(1, cst["rays_shadow_env"] + cst["rays_shadow_light"])
        shadow_nodes_per_ray = (cst["nodes_visited"] - cst["nodes_visited_extend"]) / sh_rays_c
        shadow_tris_per_ray = (cst["tris_tested"] - cst["tris_tested_extend"]) / sh_rays_c
        bytes_per_ray = 80.0 * nodes_per_ray + 48.0 * tris_per_ray + 48.0
        ext_rays = stage["rays_extension"]
        ext_s = stage["extend_ms"] * 1e-3
        n_launch = max(1, int(stage["extend_launches"]))
        achieved = bytes_per_ray * ext_rays / ext_s / 1e9 if ext_s > 0 else 0.0
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        issue_peak = n_sm * 4 * sm_hz
        traffic, issue, memory, kernels = None, None, None, None
        if ncu is not None:
            # everything below is per launch of k_extend, like `achieved`: the wave the ncu subprocess replayed is the wave a step renders
            ke = ncu["k_extend"]
            traffic = ke["dram_bytes"] / ke["launches"]
            ach = ke["warp_inst"] / ext_s * (n_launch / ke["launches"])
            tpi = ke["thread_inst"] / max(1.0, ke["warp_inst"])
            issue = {"achieved": ach, "peak": issue_peak, "unit": "warp-inst/s", "frac": ach / issue_peak,
                     "warp_instructions_per_launch": ke["warp_inst"] / ke["launches"], "warp_instructions_per_ray": ke["warp_inst"] / max(1, ext_rays) * (n_launch / ke["launches"]),
                     "simt_threads_per_inst": tpi, "useful_lane_frac": ach / issue_peak * tpi / 32.0,
                     "source": ncu_src + " / live CUDA-event launch time; peak = %d SMs x 4 schedulers x %.0f MHz (sampled during the timed region)" % (n_sm, sm_hz / 1e6)}
            memory = {"dram_gbs": ke["dram_bytes"] / ext_s / 1e9 * (n_launch / ke["launches"]), "l2_gbs": ke["l2_bytes"] / ext_s / 1e9 * (n_launch / ke["launches"]),
                      "dram_frac_of_peak": ke["dram_bytes"] / ext_s / 1e9 * (n_launch / ke["launches"]) / peak,
                      "dram_bytes_per_ray": ke["dram_bytes"] / max(1, ext_rays) * (n_launch / ke["launches"]), "l2_bytes_per_ray": ke["l2_bytes"] / max(1, ext_rays) * (n_launch / ke["launches"]),
                      "note": "ncu dram__bytes_{read,write}.sum and 32 B x lts__t_sectors_op_{read,write}.sum of the k_extend launches of one wave / their live CUDA-event time"}
            kernels = {k: {"launches": v["launches"], "warp_inst": v["warp_inst"], "threads_per_inst": v["thread_inst"] / max(1.0, v["warp_inst"]),
                           "dram_bytes": v["dram_bytes"]} for k, v in ncu.items()}
        else:
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            try:
                tj = json.load(open(tp))
                stale = tj.get("git") != sha or tj.get("workload") != args.workload
                traffic = tj.get("k_extend_dram_bytes_per_launch")
                wi = tj.get("k_extend_warp_instructions_per_launch")
                if wi and ext_s > 0:
                    ach = wi * n_launch / ext_s
                    issue = {"achieved": ach, "peak": issue_peak, "unit": "warp-inst/s", "frac": ach / issue_peak, "warp_instructions_per_launch": wi,
                             "stale": bool(stale),
                             "source": "FALLBACK (%s): committed profiles/traffic.json (captured at git %s, workload %s; this run is git %s, workload %s%s) / live launch time"
                                       % (ncu_src, tj.get("git"), tj.get("workload"), sha, args.workload, " - STALE, do not quote" if stale else "")}
            except Exception as e:       # noqa: BLE001
                print("bench.py: profiles/traffic.json unusable: %r" % (e,), file=sys.stderr)
        if args.workload == "clock":
            wl = "ClockCC0 stand-in (125281 tris, 12x%d^2 8-bit maps, 4096x2048 HDRI, defocus) %dx%d, %d spp/step/GPU, mode=%s" % (args.tex, W, H, S_, args.mode)
            l2 = "inputs larger than L2 (textures 805 MB + HDRI 134 MB + 12 GB wave state vs 126 MB L2); no flush needed"
            note = ("BVH8 + triangles of this scene (~8 MB) are L2-resident: the HBM roof is a loose upper bound for this kernel, which is bound by instruction issue "
                    "(see 'issue': warp instructions and threads/instruction measured by ncu in this run); 'memory' = the bytes it really moves")
        else:
            wl = "displaced grid (BASELINE configs[3]: %d triangles, sun+sky 4096x2048 HDRI) %dx%d, %d spp/step/GPU, mode=%s" % (len(sc.tris), W, H, S_, args.mode)
            l2 = "BVH8 nodes + triangle slots + shading triangles (%.0f MB) and 12 GB of wave state vs 126 MB L2; no flush needed" % ((st["bvh_nodes"] * 80 + st["bvh_tri_slots"] * 48 + len(sc.tris) * 144) / 1e6)
            note = "node + triangle fetches of this scene miss L2: 'memory' holds the DRAM / L2 bytes per ray and GB/s ncu measured in this run against the HBM peak"
        out = {
            "metric": "samples_per_second", "value": value, "unit": "pixel-samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "ms_per_step_host": dt_host / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "git": sha,
                       "spp_per_step": S_, "parallelism": "sample-split x%d, scene replicated, 1 native ncclReduce of the film records (sum.xyz + count) to rank 0 per step" % world,
                       "l2": l2,
                       "wave_spp": "auto (16 samples of every pixel in flight per wave)" if (args.wave_spp == 0 and args.mode == "fast") else (args.wave_spp if args.mode == "fast" else 1),
                       "timing": "value: CUDA events on the context's stream around every eleven_render and every eleven_reduce_film (ncclReduce on the same stream), max over ranks, "
                                 "bracketed by barrier + cudaDeviceSynchronize; ms_per_step_host and e2e: host clock between the same synchronisations"},
            "mrays_per_s": rays_total / dt / 1e6,
            "frame_spp_per_s": S_ * args.steps * world / dt,
            "kpaths_per_s_rank0": st["hit_bounces"] / (dt * 1e3),          # the reference's own unit: hit bounces per millisecond (S/main.cpp:172-179), this rank
            "e2e": {"value": total / dt_e2e, "unit": "pixel-samples/s", "h2d_bytes_per_step": 56, "d2h_bytes_per_step": W * H * 16},
            "gpu_launches": int(st["kernel_launches"] - launches0),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "k_extend (closest hit over BVH8)", "peak_source": peak_src, "issue": issue, "memory": memory, "kernels": kernels,
                         "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray, "exact_key_evals_per_ray": keys_per_ray,
                         "shadow_nodes_per_ray": shadow_nodes_per_ray, "shadow_tris_per_ray": shadow_tris_per_ray,
                         "launch_ms": stage["extend_ms"] / max(1, stage["extend_launches"]), "launches": int(stage["extend_launches"]),
                         "share_of_step": stage["extend_ms"] / stage["render_ms"] if stage["render_ms"] else None,
                         "stage_ms": {k: stage[k] for k in ("extend_ms", "shade_ms", "connect_ms", "other_ms", "render_ms")},
                         "reduce_ms_per_step": (st["reduce_ms"] - st0["reduce_ms"]) / args.steps, "note": note},
            "setup": {"scene_upload_s": setup_s, "bvh_build_ms": st["bvh_build_ms"], "bvh_nodes": st["bvh_nodes"], "key_slack": st["key_slack"]},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, flat)
        print(json.dumps(out), file=json_out, flush=True)
    r.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
