#!/bin/bash
# Quick multi-GPU check (gpurun --gpus 2): native reduce tests, torchrun bench at N = 2, CLI job at N = 1, 2.
tag=${1:-mc}; out=gpurun_out/$tag; mkdir -p $out
export PYTHONUNBUFFERED=1
python -c "import bench, argparse; a=argparse.Namespace(tex=4096,width=1920,height=1080,workload='clock',grid=0); print(bench.get_scene(a, need_dir=False))" > $out/scene.log 2>&1
timeout 600 python -m pytest tests/test_gpu_modes.py -q -k "two_gpus or native_film_reduce" > $out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-ncu > $out/bench_2gpu.json 2> $out/bench_2gpu.err; echo "bench 2 rc=$?"
python -c "import json; d=json.load(open('$out/bench_2gpu.json')); print('N=2 value %.1f M/s e2e %.1f ms/step %.2f reduce %.3f ms' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['roofline']['reduce_ms_per_step']))"
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $out/bench_ref_n2_rank0only.json 2> /dev/null; cut -c1-200 $out/bench_ref_n2_rank0only.json
for n in 1 2; do tfg-pathtracer_b200/host/eleven /tmp/eleven_bench_cache/clock_t4096_1920x1080.flat 1000 /tmp/o_$n.bmp --gpus $n > $out/job_g$n.log 2>&1; grep "job:" $out/job_g$n.log; done
