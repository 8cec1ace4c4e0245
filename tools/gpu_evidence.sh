#!/bin/bash
# Evidence session: grid-workload bench (config 4: the HBM-bound case) with live ncu counters, ncu launch list of bench.py itself,
# ncu --set full of the three top kernels, then the converged-reference fixtures (long: the reference at 16 384 spp).
# usage (under gpurun, from the repo root): bash tools/gpu_evidence.sh <tag>
tag=${1:-ev}
out=gpurun_out/$tag
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python bench.py --workload grid --no-cpu-baseline > $out/bench_grid.json 2> $out/bench_grid.err; echo "bench grid rc=$?"; tail -c 1500 $out/bench_grid.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ncu > $out/launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shadowEnv" -c 6 -o $out/trace_full -f python tools/profile_run.py --spp 16 > $out/ncu_trace.log 2>&1; echo "ncu trace rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_shade$" -c 2 -o $out/shade_full -f python tools/profile_run.py --spp 16 > $out/ncu_shade.log 2>&1; echo "ncu shade rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"k_extend" -c 2 -o $out/grid_extend_full -f python tools/profile_run.py --spp 16 --workload grid > $out/ncu_grid.log 2>&1; echo "ncu grid rc=$?"
if [ -z "$SKIP_CONV" ]; then
timeout 3000 python tests/golden/make_convergence.py --fullframe > $out/make_convergence.log 2>&1; echo "convergence rc=$?"; tail -c 1200 $out/make_convergence.log
fi
ls -la $out gpurun_out/*.npz
