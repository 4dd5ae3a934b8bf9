#!/bin/bash
# N-GPU box: bit-identity tests + strong-scaling bench lines (short). Usage: bash tools/gpu_multi4.sh <tag> <N>
TAG=${1:-multi4}
N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q > $OUT/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multigpu.log
tail -n 4 $OUT/pytest_multigpu.log
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_n1.json 2> $OUT/bench_c2_n1.err
for n in 2 $N; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_n$n.json 2> $OUT/bench_c2_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --workload c4 > $OUT/bench_c4_n$N.json 2> $OUT/bench_c4_n$N.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --workload c3s > $OUT/bench_c3s_n$N.json 2> $OUT/bench_c3s_n$N.err
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "n", d["n_gpus"], round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-1200:])
PY
done
rm -rf gpurun_out/scenes
