#!/bin/bash
# Job-level strong scaling only: `eleven <flat> <spp> out.bmp --gpus n` for n = 1, 2, 4, 8 (as many as the box has), twice each.
# usage (gpurun --gpus 8): bash tools/gpu_jobs.sh <tag> [spp]
tag=${1:-jobs}; spp=${2:-1000}
out=gpurun_out/$tag
mkdir -p $out
ngpu=$(nvidia-smi -L | wc -l)
python -c "import bench, argparse; a=argparse.Namespace(tex=4096,width=1920,height=1080,workload='clock',grid=0); print(bench.get_scene(a, need_dir=False))" > $out/scene.log 2>&1
flat=/tmp/eleven_bench_cache/clock_t4096_1920x1080.flat
exe=tfg-pathtracer_b200/host/eleven
for n in 1 2 4 8; do
  [ $n -le $ngpu ] || continue
  for rep in 1 2; do
    $exe $flat $spp /tmp/o_$n.bmp --gpus $n > $out/job_flat_g${n}_r$rep.log 2>&1; echo "n=$n rep=$rep $(grep 'job:' $out/job_flat_g${n}_r$rep.log)"
  done
done
python - <<PY
import numpy as np
a=np.fromfile("/tmp/o_1.bmp",np.uint8)[54:].astype(int); b=np.fromfile("/tmp/o_$ngpu.bmp",np.uint8)[54:].astype(int)
print("BMP 1 vs $ngpu GPUs: bytes differing %.6f, max |d| %d" % ((a!=b).mean(), np.abs(a-b).max()))
PY
