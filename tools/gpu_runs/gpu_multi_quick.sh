#!/bin/bash
# N-GPU box, short: the ReSTIR DI workloads at N ranks (reservoir exchange inside render) -- bench lines only
TAG=${1:-multiq}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for w in c4 c5; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --workload $w > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "n", d["n_gpus"], round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), d.get("pass_ms_per_step"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
rm -rf gpurun_out/scenes
