#!/bin/bash
# full GPU suite, the way the driver runs it
OUT=gpurun_out/${1:-tests}
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -n 40 $OUT/pytest_gpu.log
rm -rf gpurun_out/scenes
