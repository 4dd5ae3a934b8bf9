#!/bin/bash
# 640x360 parity of the ReSTIR DI variants that only had 96x54 fixtures: light presampling, directional reuse, split passes
OUT=gpurun_out/parity_restir
mkdir -p $OUT
C="timeout 400 python tools/gpu_check.py --reps 1 --width 640 --height 360 --spp 2 --bounces 4"
$C --variant c4p --scene synthetic_20k --envmap data/sky_128x64.b2env --fused 1 --tag _final_c4p_synenv > $OUT/c4p.log 2>&1
$C --variant c4 --scene cornell_pbr_16x9 --fused 1 --directional --tag _final_cornell_directional > $OUT/dir.log 2>&1
$C --variant c4 --scene cornell_pbr_16x9 --fused 0 --tag _final_cornell_split > $OUT/split.log 2>&1
$C --variant c4p --scene cornell_pbr_16x9 --fused 0 --tag _final_c4p_cornell_split > $OUT/c4p_split.log 2>&1
tail -n 3 $OUT/*.log | cut -c 1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/gpu_check_c4*_final_*.json")):
    d=json.load(open(f))
    for k in ("cuda_precise_vs_ref","cuda_fast_vs_ref","port_vs_ref"):
        if k in d:
            r=d[k]; print(f.split('/')[-1][10:], k, "over", r['pixels_over_tol'], "/", r['pixels'], "bit-exact", r['pixels_bit_exact'], "prim", r['prim_mismatch'])
PY
