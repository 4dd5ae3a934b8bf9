#!/bin/bash
# 10M-triangle synthetic scene on the GPU box: parity on an oracle-rendered window + bench lines + ncu.
TAG=${1:-syn}
OUT=gpurun_out/$TAG
mkdir -p $OUT gpurun_out/scenes
python tools/make_synthetic_scene.py gpurun_out/scenes/synthetic_10000000_100000_1920x1080.b2scene > $OUT/scene_info.json
python tools/make_envmap.py gpurun_out/scenes/sky_2048x1024.b2env
S=gpurun_out/scenes/synthetic_10000000_100000_1920x1080.b2scene
E=gpurun_out/scenes/sky_2048x1024.b2env
W="--window 896 476 1024 604"
timeout 900 python tools/gpu_check.py --variant c3 --scene $S --width 1920 --height 1080 --spp 1 --bounces 8 $W --reps 3 --tag _syn10m > $OUT/check_c3.log 2>&1
timeout 900 python tools/gpu_check.py --variant c4 --scene $S --envmap $E --width 1920 --height 1080 --spp 2 --bounces 8 $W --reps 2 --fused 1 --tag _syn10m_env > $OUT/check_c4.log 2>&1
for w in c3s c4s; do
  timeout 900 python bench.py --steps 2 --warmup 3 --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_persistent -s 3 -c 1 -o $OUT/c3s_megakernel -f python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload c3s > $OUT/ncu_full.log 2>&1
rm -rf gpurun_out/scenes
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/gpu_check_*syn10m*.json")):
    d=json.load(open(f))
    print(f, {k: d[k] for k in d if k.endswith("_vs_ref") or k.endswith("mpaths_per_s") or k.endswith("faults") or k.endswith("counters")})
PY
for f in $OUT/bench_*.json; do tail -c 2500 $f; echo; done
