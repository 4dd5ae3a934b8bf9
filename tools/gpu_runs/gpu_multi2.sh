#!/bin/bash
# N-GPU box: bit-identity tests of the partitioned render, then the multi-GPU bench (c5 by default)
N=${2:-2}
OUT=gpurun_out/${1:-multi}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multigpu.py tests/test_gpu_fullsize.py -m gpu -x -q -k "partitioned or rotated or hdr_envmap or update_materials" > $OUT/pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multigpu.log; tail -n 12 $OUT/pytest_multigpu.log
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${3:-3} --warmup 3 > $OUT/bench_c5_n$N.json 2> $OUT/bench_c5_n$N.err; tail -c 2500 $OUT/bench_c5_n$N.json; tail -n 8 $OUT/bench_c5_n$N.err
rm -rf gpurun_out/scenes
