#!/bin/bash
# BASELINE configs[4] as stated: 3840x2160, 4096 spp, textured Disney materials, defocus, full MIS (environment + 4 point lights) on 8 B200s,
# through the CLI (one process, 8 contexts, one native ncclReduce).  A 64-spp render on ONE GPU is the cross-check of the picture.
# usage (gpurun --gpus 8): bash tools/gpu_config5.sh <tag> [spp]
tag=${1:-c5}; spp=${2:-4096}
out=gpurun_out/$tag; mkdir -p $out
ngpu=$(nvidia-smi -L | wc -l)
python - > $out/scene.log 2>&1 <<PY
import time
from tfg_pathtracer_b200 import scenes as S
t=time.time(); sc=S.textured_lights(); S.save_flat(sc, "/tmp/config5.flat"); print("scene", len(sc.tris), sc.width, sc.height, len(sc.lights), "lights", time.time()-t, "s")
PY
exe=tfg-pathtracer_b200/host/eleven
$exe /tmp/config5.flat 64 /tmp/c5_1gpu.bmp --gpus 1 --raw /tmp/c5_1gpu.f32 > $out/job_1gpu_64spp.log 2>&1; grep "job:\|spp on" $out/job_1gpu_64spp.log
$exe /tmp/config5.flat $spp /tmp/c5_ngpu.bmp --gpus $ngpu --raw /tmp/c5_ngpu.f32 > $out/job_${ngpu}gpu_${spp}spp.log 2>&1; grep "job:\|spp on" $out/job_${ngpu}gpu_${spp}spp.log
python - <<PY | tee $out/compare.json
import numpy as np, json
a=np.fromfile("/tmp/c5_1gpu.f32",np.float32).reshape(2160,3840,4)[...,:3].astype(np.float64); b=np.fromfile("/tmp/c5_ngpu.f32",np.float32).reshape(2160,3840,4)[...,:3].astype(np.float64)
blk=lambda x: x.reshape(270,8,480,8,3).mean((1,3))
rel=np.abs(blk(a)-blk(b))/(blk(b)+1e-3)
print(json.dumps({"mean_1gpu_64spp": a.mean(), "mean_ngpu": b.mean(), "rel_mean_diff": (a.mean()-b.mean())/b.mean(), "median_block8_rel": float(np.median(rel)), "p99_block8_rel": float(np.percentile(rel,99)), "finite": bool(np.isfinite(b).all())}))
PY
