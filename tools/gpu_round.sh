#!/bin/bash
# One GPU session: parity tests, bench (both arms), A/B variants, ncu launch list + full capture of the traversal kernels.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench_ours.json 2> $out/bench_ours.err; echo "bench rc=$?"
cat $out/bench_ours.json
for v in $(ls tfg-pathtracer_b200/csrc/libeleven_b200_*.so 2>/dev/null); do
  n=$(basename $v .so)
  ELEVEN_LIB=$PWD/$v timeout 300 python bench.py --no-cpu-baseline > $out/bench_$n.json 2> $out/bench_$n.err; echo "$n rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$out/bench_$n.json")); print("$n", d["value"], d["roofline"]["stage_ms"])
except Exception as e: print("$n failed", e)
PY
done
if [ -z "$SKIP_REF" ]; then
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo "ref rc=$?"
cat $out/bench_reference.json
fi
if [ -n "$SANITIZE" ]; then
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $out/racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $out/memcheck.log
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python tools/profile_run.py --spp 16 > $out/launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shadowEnv" -c 6 -o $out/wave16 -f python tools/profile_run.py --spp 16 > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_shade$" -c 2 -o $out/shade2 -f python tools/profile_run.py --spp 16 > $out/ncu_shade.log 2>&1; echo "ncu shade rc=$?"
ls -la $out
