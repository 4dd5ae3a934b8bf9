#!/bin/bash
# A/B bench of experiment builds (libb2pt_<tag>.so) on the C2 workload. Usage: bash tools/gpu_ab.sh <outtag> [workload]
TAG=${1:-ab}
WL=${2:-c2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $WL"
$B > $OUT/bench_default.json 2> $OUT/bench_default.err
for lib in hiprt-path-tracer_b200/libb2pt_*.so; do
  t=$(basename $lib .so); t=${t#libb2pt_}
  B2PT_LIB=$lib $B > $OUT/bench_$t.json 2> $OUT/bench_$t.err
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "mk_ms", d.get("pass_ms_per_step"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-500:])
PY
done
