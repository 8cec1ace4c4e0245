#!/bin/bash
# A/B of triangle pre-splitting (ELEVEN_PRESPLIT=0/1) + the GPU suite on the new trees.  usage: bash tools/gpu_presplit.sh <tag>
tag=${1:-ps}
out=gpurun_out/$tag
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $out/pytest_gpu.log
short() {
  n=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-ncu > $out/bench_$n.json 2> $out/bench_$n.err
  python - <<PY
import json
try:
    d = json.load(open("$out/bench_$n.json")); s = d["roofline"]["stage_ms"]; r = d["roofline"]
    print("%-12s %8.1f M/s step %.2f ms | ext %.2f shade %.2f conn %.2f | nodes/ray %.2f tris/ray %.2f shadow %.2f/%.2f | bvh %d nodes %.1f ms" % ("$n", d["value"] / 1e6, d["ms_per_step"], s["extend_ms"], s["shade_ms"], s["connect_ms"], r["nodes_per_ray"], r["tris_per_ray"], r["shadow_nodes_per_ray"], r["shadow_tris_per_ray"], d["setup"]["bvh_nodes"], d["setup"]["bvh_build_ms"]))
except Exception as e: print("$n failed", e)
PY
}
short presplit1 ELEVEN_PRESPLIT=1
short presplit0 ELEVEN_PRESPLIT=0
short presplit1b ELEVEN_PRESPLIT=1
short ps_a1s4 ELEVEN_PRESPLIT_AREA=1.0 ELEVEN_PRESPLIT_SLIVER=4 ELEVEN_PRESPLIT_DEPTH=6
short ps_a2s4 ELEVEN_PRESPLIT_AREA=2.0 ELEVEN_PRESPLIT_SLIVER=4 ELEVEN_PRESPLIT_DEPTH=6
short ps_a4s8 ELEVEN_PRESPLIT_AREA=4.0 ELEVEN_PRESPLIT_SLIVER=8 ELEVEN_PRESPLIT_DEPTH=5
timeout 300 python bench.py --workload grid --no-cpu-baseline --no-ncu > $out/bench_grid.json 2> $out/bench_grid.err; python -c "import json; d=json.load(open('$out/bench_grid.json')); print('grid %.1f M/s upload %.3f s bvh %.1f ms' % (d['value']/1e6, d['setup']['scene_upload_s'], d['setup']['bvh_build_ms']))"
