#!/bin/bash
# r02: precise-flavour optimisation round: parity first, then A/B, then the new bench defaults
OUT=gpurun_out/${1:-r02b}
mkdir -p $OUT
timeout 120 ./tools/div_selftest 20000 > $OUT/div_selftest.json 2>&1; echo "div_selftest rc=$?"; cat $OUT/div_selftest.json
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "rc=$?" >> $OUT/pytest_parity.log; tail -n 8 $OUT/pytest_parity.log
timeout 900 python tools/flavour_ab.py --flavours precise,fast --tag main --workloads c2,c3s,c4s > $OUT/ab_main.json 2> $OUT/ab_main.err; cat $OUT/ab_main.json; tail -n 3 $OUT/ab_main.err
if [ "$2" != "nobench" ]; then
timeout 1500 python bench.py --steps 3 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 3000 $OUT/bench_default.json; tail -n 5 $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json; tail -n 5 $OUT/bench_reference.err
fi
if [ "$3" == "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_phased -s 3 -c 1 -o $OUT/c2_precise -f python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload c2 > $OUT/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_phased -s 3 -c 1 -o $OUT/c3s_precise -f python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload c3s > $OUT/ncu_c3s.log 2>&1
tail -n 2 $OUT/ncu_c2.log $OUT/ncu_c3s.log | cut -c1-300
fi
rm -rf gpurun_out/scenes
