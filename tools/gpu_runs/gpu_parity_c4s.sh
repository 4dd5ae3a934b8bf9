#!/bin/bash
# BASELINE configs[3] parity at full size after the wide-evaluation change: 10M triangles + 2k envmap, ReSTIR DI fused, 2 frames,
# oracle window 256x256 at the frame centre, compared on its interior (margin 24 px > reuse radius 16 px x 1 reuse frame)
OUT=gpurun_out/parity_c4s
mkdir -p $OUT gpurun_out/scenes
S=gpurun_out/scenes/s10m.b2scene
E=gpurun_out/scenes/sky_2048x1024.b2env
python tools/make_synthetic_scene.py $S > $OUT/scene_info.json
python tools/make_envmap.py $E
timeout 500 python tools/gpu_check.py --variant c4 --scene $S --envmap $E --width 1920 --height 1080 --spp 2 --bounces 8 --window 832 412 1088 668 --margin 24 --reps 1 --fused 1 --tag _final_syn10m_env_margin > $OUT/check_c4.log 2>&1
tail -n 3 $OUT/check_c4.log
rm -rf gpurun_out/scenes
python - <<'PY'
import json
d=json.load(open("gpurun_out/gpu_check_c4_1920x1080_2spp_final_syn10m_env_margin.json"))
for k in ("cuda_precise_vs_ref","cuda_fast_vs_ref","port_vs_ref","cuda_precise_vs_port"):
    r=d[k]; print(k, "over", r['pixels_over_tol'], "/", r['pixels'], "bit-exact", r['pixels_bit_exact'], "prim", r['prim_mismatch'])
PY
