#!/bin/bash
# per-kernel wavefront times (CUDA events) + ncu --set full of round 1 and round 4 of the c3s MegaKernel pass
OUT=gpurun_out/${1:-r02p}
mkdir -p $OUT
timeout 900 python tools/flavour_ab.py --flavours precise --tag main --workloads c3s,c4s > $OUT/ab_main.json 2> $OUT/ab_main.err; cat $OUT/ab_main.json; tail -n 3 $OUT/ab_main.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"wf_trace|wf_stage1|wf_stage2" -s 3 -c 4 -o $OUT/c3s_wavefront_round1 -f python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline --workload c3s > $OUT/ncu_c3s_r1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"wf_trace|wf_stage1|wf_stage2" -s 15 -c 4 -o $OUT/c3s_wavefront_round4 -f python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline --workload c3s > $OUT/ncu_c3s_r4.log 2>&1
tail -n 2 $OUT/ncu_c3s_r1.log $OUT/ncu_c3s_r4.log | cut -c1-200
rm -rf gpurun_out/scenes
