#!/bin/bash
# compute-sanitizer over tools/sanitize_run.py (both BVH builders incl. pre-split references, closest-hit batch, parity + fast renders on two
# streams, native reduce, the material zoo with a point light).  usage: bash tools/gpu_sanitize.sh <tag>
tag=${1:-san}
out=gpurun_out/$tag
mkdir -p $out
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $out/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $out/racecheck.log
