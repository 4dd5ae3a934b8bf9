#!/bin/bash
# ncu launch list (per-kernel durations, cold cache, serialised) of one bench step: tools/gpu_launchlist.sh TAG WORKLOAD
OUT=gpurun_out/$1
mkdir -p $OUT
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active --clock-control none -c 2000 --csv --log-file $OUT/launches_$2.csv python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload $2 > $OUT/launches_$2.log 2>&1
tail -n 2 $OUT/launches_$2.log | cut -c1-200
wc -l $OUT/launches_$2.csv
rm -rf gpurun_out/scenes
