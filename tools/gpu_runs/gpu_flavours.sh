#!/bin/bash
# r02: attribute the precise/fast gap to compiler flags; ncu source-level captures of the precise MegaKernel
OUT=gpurun_out/r02a
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_parity.log 2>&1; echo "rc=$?" >> $OUT/pytest_parity.log; tail -n 5 $OUT/pytest_parity.log
timeout 900 python tools/flavour_ab.py --flavours precise,fast --tag main > $OUT/ab_main.json 2> $OUT/ab_main.err; cat $OUT/ab_main.json
for lib in hiprt-path-tracer_b200/libb2pt_p*.so; do
  tag=$(basename $lib .so)
  B2PT_LIB=$lib timeout 600 python tools/flavour_ab.py --flavours precise --tag $tag > $OUT/ab_$tag.json 2> $OUT/ab_$tag.err; cat $OUT/ab_$tag.json; tail -n 3 $OUT/ab_$tag.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_phased -s 3 -c 1 -o $OUT/c2_precise -f python bench.py --precise 1 --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload c2 > $OUT/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_phased -s 3 -c 1 -o $OUT/c3s_precise -f python bench.py --precise 1 --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload c3s > $OUT/ncu_c3s.log 2>&1
tail -n 3 $OUT/ncu_c2.log $OUT/ncu_c3s.log
rm -rf gpurun_out/scenes
