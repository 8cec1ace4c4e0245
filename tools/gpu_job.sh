#!/bin/bash
# Job-level timing of the CLI as a user runs it (process start -> BMP on disk): `eleven <scene> <spp> out.bmp --gpus N` for N = 1 and
# every power of two up to the GPUs of the box, on the bench scene (flat container and reference scene directory).
# usage (under gpurun, from the repo root): bash tools/gpu_job.sh <tag> [spp]
tag=${1:-job}; spp=${2:-1000}
out=gpurun_out/$tag
mkdir -p $out
python -c "import bench, argparse; a=argparse.Namespace(tex=4096,width=1920,height=1080,workload='clock',grid=0); print(bench.get_scene(a, need_dir=True))" > $out/scene.log 2>&1
flat=/tmp/eleven_bench_cache/clock_t4096_1920x1080.flat
dir=/tmp/eleven_bench_cache/clock_t4096_1920x1080_dir
exe=tfg-pathtracer_b200/host/eleven
ngpu=$(nvidia-smi -L | wc -l)
( cd $dir && ELEVEN_UPLOAD_TRACE=1 $OLDPWD/$exe $dir 64 /tmp/o_dir.bmp ) > $out/job_dir_trace.log 2>&1; grep -E "job:|eleven_scene_upload|loaded" $out/job_dir_trace.log
for n in 1 2 4 8; do
  [ $n -le $ngpu ] || continue
  for rep in 1 2; do
    $exe $flat $spp /tmp/o_$n.bmp --gpus $n > $out/job_flat_g${n}_r$rep.log 2>&1; grep "job:" $out/job_flat_g${n}_r$rep.log
  done
done
