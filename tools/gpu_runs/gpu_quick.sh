#!/bin/bash
# parity (bounded) + one bench line per workload with the default build
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden and c1_128" > $OUT/pytest_first.log 2>&1
rc=$?; tail -n 3 $OUT/pytest_first.log
if [ $rc -eq 124 ]; then echo "HANG in first test; stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 30 $OUT/pytest_gpu.log
for w in c1 c2 c3 c4 c3s c4s; do
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err
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
rm -rf gpurun_out/scenes
