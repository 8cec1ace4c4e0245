#!/bin/bash
# A/B session: GPU tests on the main library, the full bench, then one short bench per variant library / environment knob.
# usage (under gpurun, from the repo root): bash tools/gpu_ab.sh <tag> [skiptests]
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
export PYTHONUNBUFFERED=1
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest_gpu.log
fi
timeout 600 python bench.py > $out/bench_main.json 2> $out/bench_main.err; echo "bench rc=$?"
short() {   # name, env assignments...
  n=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-ncu > $out/bench_$n.json 2> $out/bench_$n.err
  python - <<PY
import json
try:
    d = json.load(open("$out/bench_$n.json")); s = d["roofline"]["stage_ms"]
    print("%-14s %8.1f M/s  e2e %8.1f  step %.2f ms | ext %.2f shade %.2f conn %.2f other %.2f" % ("$n", d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], s["extend_ms"], s["shade_ms"], s["connect_ms"], s["other_ms"]))
except Exception as e: print("$n failed", e)
PY
}
short main X=1
short nooverlap ELEVEN_OVERLAP=0
for v in $(ls tfg-pathtracer_b200/csrc/libeleven_b200_*.so 2>/dev/null); do
  n=$(basename $v .so); n=${n#libeleven_b200_}
  short $n ELEVEN_LIB=$PWD/$v
done
short main2 X=1
ls $out
