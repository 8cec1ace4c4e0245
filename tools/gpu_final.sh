#!/bin/bash
# Final evidence session of a round on one B200: the -m gpu suite, both bench arms, the ncu launch list of bench.py itself, ncu --set full of
# the three top kernels (with source, for tools/ncu_regions.py), the CLI job from the reference-format scene directory.
# usage (under gpurun, from the repo root): bash tools/gpu_final.sh <tag>
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
export PYTHONUNBUFFERED=1
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.log; fi
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo "ref rc=$?"; cut -c1-400 $out/bench_reference.json
timeout 600 python bench.py > $out/bench_ours.json 2> $out/bench_ours.err; echo "bench rc=$?"; cut -c1-300 $out/bench_ours.json
cp gpurun_out/bench_ncu_counters_clock.csv $out/ 2>/dev/null
timeout 300 python bench.py --mode parity --no-ncu --no-cpu-baseline > $out/bench_parity.json 2> $out/bench_parity.err; echo "parity rc=$?"
timeout 300 python bench.py --hit-mode min_t --no-ncu --no-cpu-baseline > $out/bench_min_t.json 2> $out/bench_min_t.err; echo "min_t rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ncu > $out/launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shadowEnv" -c 4 -o $out/trace_full -f python tools/profile_run.py --spp 16 > $out/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_shade$" -c 2 -o $out/shade_full -f python tools/profile_run.py --spp 16 > $out/ncu_shade.log 2>&1; echo "ncu shade rc=$?"
dir=/tmp/eleven_bench_cache/clock_t4096_1920x1080_dir
( cd $dir && ELEVEN_UPLOAD_TRACE=1 $OLDPWD/tfg-pathtracer_b200/host/eleven $dir 1000 /tmp/o_dir.bmp ) > $out/job_dir_1gpu.log 2>&1; grep -E "job:|eleven_scene_upload|loaded" $out/job_dir_1gpu.log
ls $out
for v in $(ls tfg-pathtracer_b200/csrc/libeleven_b200_*.so 2>/dev/null); do
  n=$(basename $v .so); n=${n#libeleven_b200_}
  ELEVEN_LIB=$PWD/$v timeout 300 python bench.py --no-cpu-baseline --no-ncu > $out/bench_var_$n.json 2> $out/bench_var_$n.err
  python -c "import json; d=json.load(open('$out/bench_var_$n.json')); s=d['roofline']['stage_ms']; print('variant $n %.1f M/s ext %.2f conn %.2f' % (d['value']/1e6, s['extend_ms'], s['connect_ms']))"
done
timeout 300 python bench.py --no-cpu-baseline --no-ncu > $out/bench_again.json 2> $out/bench_again.err; python -c "import json; d=json.load(open('$out/bench_again.json')); s=d['roofline']['stage_ms']; print('main again %.1f M/s ext %.2f conn %.2f' % (d['value']/1e6, s['extend_ms'], s['connect_ms']))"
