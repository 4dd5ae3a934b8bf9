#!/bin/bash
# phased MegaKernel: parity first (bounded), then A/B benches against the other two forms
TAG=${1:-ph}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden and c1_128" > $OUT/pytest_first.log 2>&1
rc=$?; echo "first rc=$rc" >> $OUT/pytest_first.log; tail -n 15 $OUT/pytest_first.log
if [ $rc -eq 124 ]; then echo "HANG in first test; stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 25 $OUT/pytest_gpu.log
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
for form in phased persistent tiles; do
  B2PT_MEGAKERNEL=$form $B > $OUT/bench_c2_$form.json 2> $OUT/bench_c2_$form.err
done
for w in c1 c3 c4; do
  $B --workload $w > $OUT/bench_${w}_phased.json 2> $OUT/bench_${w}_phased.err
done
for form in phased persistent; do
  B2PT_MEGAKERNEL=$form timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c3s > $OUT/bench_c3s_$form.json 2> $OUT/bench_c3s_$form.err
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "mk_ms", d.get("pass_ms_per_step"), d.get("traversal"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_phased -s 3 -c 1 -o $OUT/c2_megakernel -f python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
rm -rf gpurun_out/scenes
