#!/bin/bash
# Second GPU session of a round: DRAM traffic of all k_extend launches of one wave (ncu --set full), the full-size configs 2/4/5
# with their parity properties, the BVH builder benchmark.  usage: bash tools/gpu_round2.sh <tag>
tag=${1:-run2}
out=gpurun_out/$tag
mkdir -p $out
export PYTHONUNBUFFERED=1
python -c "import bench, argparse; a=argparse.Namespace(tex=4096,width=1920,height=1080); bench.get_scene(a, need_dir=False)" > $out/scene.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_extend" -c 5 -o $out/extend5 -f python tools/profile_run.py --spp 16 > $out/ncu_extend.log 2>&1; echo "ncu extend rc=$?"
for c in 2 4 5; do
  extra=""; [ $c = 5 ] && extra="--spp 128"
  timeout 900 python tests/fullsize_config.py --config $c $extra > $out/config_$c.json 2> $out/config_$c.err; echo "config $c rc=$?"; tail -c 600 $out/config_$c.json; echo
done
timeout 600 python tools/bvh_build_bench.py > $out/bvh_build.jsonl 2> $out/bvh_build.err; echo "bvh bench rc=$?"; tail -4 $out/bvh_build.jsonl
ls -la $out
