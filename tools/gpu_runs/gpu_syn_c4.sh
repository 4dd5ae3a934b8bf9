#!/bin/bash
# C4 (ReSTIR DI fused + envmap) on the 10M-triangle scene: parity on the interior of an oracle-rendered window
TAG=${1:-syn_c4}
OUT=gpurun_out/$TAG
mkdir -p $OUT gpurun_out/scenes
python tools/make_synthetic_scene.py gpurun_out/scenes/s10m.b2scene > $OUT/scene_info.json
python tools/make_envmap.py gpurun_out/scenes/sky_2048x1024.b2env
S=gpurun_out/scenes/s10m.b2scene
E=gpurun_out/scenes/sky_2048x1024.b2env
timeout 900 python tools/gpu_check.py --variant c4 --scene $S --envmap $E --width 1920 --height 1080 --spp 2 --bounces 8 --window 832 412 1088 668 --margin 24 --reps 1 --fused 1 --tag _syn10m_env_margin > $OUT/check_c4.log 2>&1
rm -rf gpurun_out/scenes
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/gpu_check_c4*margin*.json")):
    d=json.load(open(f))
    print(f, {k: d[k] for k in d if k.endswith("_vs_ref") or k.endswith("faults")})
PY
tail -n 3 $OUT/check_c4.log | cut -c1-300
