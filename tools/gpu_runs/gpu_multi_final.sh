#!/bin/bash
# 2-GPU box, short: the multi-GPU bit-identity tests (incl. light presampling) + the default bench line at N ranks
TAG=${1:-multif}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q > $OUT/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multigpu.log
tail -n 6 $OUT/pytest_multigpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_default_n$N.json 2> $OUT/bench_default_n$N.err
tail -n 1 $OUT/bench_default_n$N.json | cut -c 1-600
rm -rf gpurun_out/scenes
