#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the 2-GPU native-reduce test, bench.py under torchrun at every power of two up to the box's GPUs,
# and the CLI job (`eleven <flat> 1000 out.bmp --gpus n`) timed from process start to BMP.
# usage: bash tools/gpu_multi.sh <tag> [spp]
tag=${1:-multi}; spp=${2:-1000}
out=gpurun_out/$tag
mkdir -p $out
export PYTHONUNBUFFERED=1
ngpu=$(nvidia-smi -L | wc -l)
python -c "import bench, argparse; a=argparse.Namespace(tex=4096,width=1920,height=1080,workload='clock',grid=0); print(bench.get_scene(a, need_dir=False))" > $out/scene.log 2>&1
timeout 600 python -m pytest tests/test_gpu_modes.py -q -k "two_gpus or native_film_reduce" > $out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_multi.log
flat=/tmp/eleven_bench_cache/clock_t4096_1920x1080.flat
exe=tfg-pathtracer_b200/host/eleven
for n in 1 2 4 8; do
  [ $n -le $ngpu ] || continue
  for rep in 1 2; do
    $exe $flat $spp /tmp/o_$n.bmp --gpus $n > $out/job_flat_g${n}_r$rep.log 2>&1; grep "job:" $out/job_flat_g${n}_r$rep.log
  done
done
cmp /tmp/o_1.bmp /tmp/o_$ngpu.bmp && echo "BMP of 1 and $ngpu GPUs identical" || python - <<PY
import numpy as np
a=np.fromfile("/tmp/o_1.bmp",np.uint8)[54:].astype(int); b=np.fromfile("/tmp/o_$ngpu.bmp",np.uint8)[54:].astype(int)
print("BMP 1 vs $ngpu GPUs: bytes differing %.5f, max |d| %d" % ((a!=b).mean(), np.abs(a-b).max()))
PY
for n in 2 4 8; do
  [ $n -le $ngpu ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 4 --warmup 3 --no-ncu > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err; echo "bench $n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_${n}gpu.json")); print("N=$n value %.1f M/s e2e %.1f ms/step %.2f reduce %.3f ms" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["roofline"]["reduce_ms_per_step"]))
except Exception as e: print("bench $n unreadable", e)
PY
done
timeout 300 python bench.py --no-ncu --no-cpu-baseline > $out/bench_1gpu.json 2> $out/bench_1gpu.err; python -c "import json; d=json.load(open('$out/bench_1gpu.json')); print('N=1 value %.1f M/s ms/step %.2f' % (d['value']/1e6, d['ms_per_step']))"
ls $out
