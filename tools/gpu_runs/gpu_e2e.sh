#!/bin/bash
# after pinning the host copies of the scene: GPU tests + the e2e numbers of the big-scene workloads and the default line
OUT=gpurun_out/e2e
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 4 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
for w in c1 c3s c4s; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],2), d["unit"], "e2e", round(d["e2e"]["value"],2), "h2d", d["e2e"]["h2d_bytes_per_step"], "e2e ms", round(d["e2e"]["ms_per_step"],1), "ms", round(d["ms_per_step"],1))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
rm -rf gpurun_out/scenes
