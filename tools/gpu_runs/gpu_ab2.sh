#!/bin/bash
# block-size A/B of the phased MegaKernel on c2 (Cornell) and c3s (10M triangles) + 2-GPU checks
TAG=${1:-ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in c2 c3s; do
  B="timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $wl"
  $B > $OUT/bench_${wl}_default.json 2> $OUT/bench_${wl}_default.err
  for lib in hiprt-path-tracer_b200/libb2pt_*.so; do
    t=$(basename $lib .so); t=${t#libb2pt_}
    B2PT_LIB=$lib $B > $OUT/bench_${wl}_$t.json 2> $OUT/bench_${wl}_$t.err
  done
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],1), d["unit"], "e2e", round(d["e2e"]["value"],1), "mk_ms", d.get("pass_ms_per_step"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
rm -rf gpurun_out/scenes
