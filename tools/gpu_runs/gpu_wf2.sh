#!/bin/bash
OUT=gpurun_out/${1:-wf2}
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -n 15 $OUT/pytest.log
timeout 900 python tools/flavour_ab.py --flavours precise,fast --tag auto --workloads ${2:-c2,c3s,c4s} > $OUT/ab_auto.json 2> $OUT/ab_auto.err; cat $OUT/ab_auto.json; tail -n 3 $OUT/ab_auto.err
rm -rf gpurun_out/scenes
