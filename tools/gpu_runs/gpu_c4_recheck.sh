#!/bin/bash
OUT=gpurun_out/c4_recheck
mkdir -p $OUT gpurun_out/scenes
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "determinism or golden" > $OUT/pytest.log 2>&1; tail -n 4 $OUT/pytest.log
C="timeout 600 python tools/gpu_check.py --variant c4 --fused 1 --reps 1 --spp 2 --bounces 4"
$C --scene synthetic_20k --envmap data/sky_128x64.b2env --width 640 --height 360 --tag _A2_res640 > $OUT/A.log 2>&1
$C --scene cornell_pbr_16x9 --width 640 --height 360 --tag _C2_cornell640 > $OUT/C.log 2>&1
timeout 600 python tools/gpu_check.py --variant c4 --fused 0 --reps 1 --spp 3 --bounces 4 --scene synthetic_20k --width 640 --height 360 --tag _F2_split > $OUT/F.log 2>&1
python tools/make_synthetic_scene.py gpurun_out/scenes/s10m.b2scene > $OUT/scene_info.json
python tools/make_envmap.py gpurun_out/scenes/sky_2048x1024.b2env
timeout 900 python tools/gpu_check.py --variant c4 --scene gpurun_out/scenes/s10m.b2scene --envmap gpurun_out/scenes/sky_2048x1024.b2env --width 1920 --height 1080 --spp 2 --bounces 8 --window 832 412 1088 668 --margin 24 --reps 1 --fused 1 --tag _syn10m_env_margin > $OUT/check_c4_10m.log 2>&1
rm -rf gpurun_out/scenes
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/gpu_check_c4_*_[ACF]2_*.json"))+sorted(glob.glob("gpurun_out/gpu_check_c4*margin*.json")):
    d=json.load(open(f))
    for k in ("cuda_precise_vs_ref","cuda_fast_vs_ref","port_vs_ref"):
        r=d[k]; print(f.split('/')[-1], k, "over", r['pixels_over_tol'], "/", r['pixels'], "prim", r['prim_mismatch'], "mean", round(r['mean_a'],4), round(r['mean_b'],4))
PY
