#!/bin/bash
OUT=gpurun_out/bisect_c4
mkdir -p $OUT gpurun_out/scenes
python tools/make_envmap.py gpurun_out/scenes/sky_2048x1024.b2env
python tools/make_envmap.py gpurun_out/scenes/sky_512x256.b2env --width 512 --height 256
E=gpurun_out/scenes/sky_2048x1024.b2env
C="timeout 600 python tools/gpu_check.py --variant c4 --fused 1 --reps 1 --spp 2 --bounces 4"
$C --scene synthetic_20k --envmap data/sky_128x64.b2env --width 640 --height 360 --tag _A_res640 > $OUT/A.log 2>&1
$C --scene synthetic_20k --envmap $E --width 256 --height 144 --tag _B_bigenv > $OUT/B.log 2>&1
$C --scene cornell_pbr_16x9 --width 640 --height 360 --tag _C_cornell640 > $OUT/C.log 2>&1
$C --scene synthetic_20k --width 256 --height 144 --tag _D_noenv > $OUT/D.log 2>&1
$C --scene synthetic_20k --envmap gpurun_out/scenes/sky_512x256.b2env --width 256 --height 144 --tag _E_env512 > $OUT/E.log 2>&1
rm -rf gpurun_out/scenes
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/gpu_check_c4_*_[A-E]_*.json")):
    d=json.load(open(f))
    for k in ("cuda_precise_vs_ref","cuda_precise_vs_port","port_vs_ref"):
        r=d[k]; print(f.split('/')[-1], k, "over", r['pixels_over_tol'], "/", r['pixels'], "prim", r['prim_mismatch'], "mean", round(r['mean_a'],4), round(r['mean_b'],4))
PY
