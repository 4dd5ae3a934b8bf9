#!/bin/bash
# End-of-round evidence at HEAD (no profiler): GPU tests, smoke, both bench arms, one line per workload.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 25 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > $OUT/smoke.log 2>&1; tail -n 2 $OUT/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
for w in c1 c3 c4 c3s c4s; do
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],2), d["unit"], "e2e", round(d["e2e"]["value"],2), "mk_ms", d.get("pass_ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
rm -rf gpurun_out/scenes
