#!/bin/bash
# ncu --set full capture of one kernel of one workload: tools/gpu_ncu.sh TAG WORKLOAD KERNEL_REGEX [precise]
OUT=gpurun_out/$1
mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$3 -s ${5:-3} -c 1 -o $OUT/$2_$3 -f python bench.py --precise ${4:-1} --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload $2 > $OUT/ncu_$2_$3.log 2>&1
tail -n 2 $OUT/ncu_$2_$3.log | cut -c1-200
rm -rf gpurun_out/scenes
