#!/bin/bash
# One GPU-box session: parity tests, A/B benches of the MegaKernel forms, ncu launch list + one full capture.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > $OUT/bench_c2_persistent.json 2> $OUT/bench_c2_persistent.err
B2PT_MEGAKERNEL=tiles $B > $OUT/bench_c2_tiles.json 2> $OUT/bench_c2_tiles.err
for t in mb8 mb10 mb12; do
  if [ -f hiprt-path-tracer_b200/libb2pt_$t.so ]; then
    B2PT_LIB=hiprt-path-tracer_b200/libb2pt_$t.so $B > $OUT/bench_c2_$t.json 2> $OUT/bench_c2_$t.err
  fi
done
for w in c1 c3 c4; do
  $B --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "mk_ms", d.get("pass_ms_per_step"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
# ncu: launch list of the bench command, then one full capture of the C2 MegaKernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c2.csv python bench.py --steps 2 --warmup 3 --spp 2 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_persistent -s 3 -c 1 -o $OUT/c2_megakernel -f python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
