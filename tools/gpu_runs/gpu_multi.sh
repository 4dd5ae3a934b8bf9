#!/bin/bash
# N-GPU box: bit-identity of the partitioned render + strong-scaling bench lines. Usage: bash tools/gpu_multi.sh <tag> <N>
TAG=${1:-multi}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 700 python -m pytest tests/test_multigpu.py -m gpu -q > $OUT/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multigpu.log
tail -n 12 $OUT/pytest_multigpu.log
run() { # workload n
  if [ "$2" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline --workload $1 > $OUT/bench_$1_n1.json 2> $OUT/bench_$1_n1.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --steps 2 --warmup 3 --no-cpu-baseline --workload $1 > $OUT/bench_$1_n$2.json 2> $OUT/bench_$1_n$2.err
  fi
}
for w in c2 c4 c3s c5; do
  run $w 1
  run $w $N
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "n", d["n_gpus"], round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
rm -rf gpurun_out/scenes
