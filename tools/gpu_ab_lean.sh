#!/bin/bash
# Lean A/B: main, every variant library, main again — short benches only.  usage: bash tools/gpu_ab_lean.sh <tag>
tag=${1:-abl}; out=gpurun_out/$tag; mkdir -p $out
short() {
  n=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-ncu > $out/bench_$n.json 2> $out/bench_$n.err
  python -c "import json; d=json.load(open('$out/bench_$n.json')); s=d['roofline']['stage_ms']; print('%-12s %8.1f M/s step %.2f ms | ext %.2f shade %.2f conn %.2f other %.2f' % ('$n', d['value']/1e6, d['ms_per_step'], s['extend_ms'], s['shade_ms'], s['connect_ms'], s['other_ms']))"
}
short main X=1
for v in $(ls tfg-pathtracer_b200/csrc/libeleven_b200_*.so 2>/dev/null); do n=$(basename $v .so); short ${n#libeleven_b200_} ELEVEN_LIB=$PWD/$v; done
short main2 X=1
