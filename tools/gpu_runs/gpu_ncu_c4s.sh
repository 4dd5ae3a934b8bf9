#!/bin/bash
# ncu --set full of the shading kernels of a c4s frame: InitialCandidates phase A, spatiotemporal phase A, stage 1 / 2 of rounds 0 and 1
OUT=gpurun_out/${1:-ncu_c4s}
mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"wf_stage1|wf_stage2|wf_pass_a" -s 0 -c 6 -o $OUT/c4s_shading -f python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline --workload c4s > $OUT/ncu_c4s.log 2>&1
tail -n 2 $OUT/ncu_c4s.log | cut -c1-200
rm -rf gpurun_out/scenes
