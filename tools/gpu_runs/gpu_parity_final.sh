#!/bin/bash
# full-size parity of the precise flavour against the reference oracle on the box's cores (after the wide-evaluation change)
OUT=gpurun_out/parity_final
mkdir -p $OUT gpurun_out/scenes
timeout 700 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -n 4 $OUT/pytest_gpu.log
C="timeout 900 python tools/gpu_check.py --reps 1"
$C --variant c1 --scene cornell_pbr_1x1 --width 512 --height 512 --spp 1 --bounces 3 --tag _final > $OUT/c1.log 2>&1
$C --variant c2 --scene cornell_pbr_16x9 --width 1920 --height 1080 --spp 2 --bounces 8 --tag _final > $OUT/c2.log 2>&1
$C --variant c3 --scene cornell_pbr_16x9 --width 1920 --height 1080 --spp 2 --bounces 8 --tag _final > $OUT/c3.log 2>&1
$C --variant c4 --scene synthetic_20k --envmap data/sky_128x64.b2env --width 640 --height 360 --spp 2 --bounces 4 --fused 1 --tag _final_synenv > $OUT/c4.log 2>&1
python tools/make_synthetic_scene.py gpurun_out/scenes/s10m.b2scene > $OUT/scene_info.json
$C --variant c3 --scene gpurun_out/scenes/s10m.b2scene --width 1920 --height 1080 --spp 1 --bounces 8 --window 896 476 1024 604 --tag _final_syn10m > $OUT/c3_10m.log 2>&1
rm -rf gpurun_out/scenes
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/gpu_check_*_final*.json")):
    d=json.load(open(f))
    for k in ("cuda_precise_vs_ref","cuda_fast_vs_ref","port_vs_ref"):
        r=d[k]; print(f.split('/')[-1][10:], k, "over", r['pixels_over_tol'], "/", r['pixels'], "bit-exact", r['pixels_bit_exact'], "prim", r['prim_mismatch'])
PY
