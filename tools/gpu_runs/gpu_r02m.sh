#!/bin/bash
# r02 evidence session: GPU parity suite, smoke, both bench arms on the default (north-star) workload, ncu launch list of the bench
# command and ncu --set full captures of the wavefront kernels on the 10M-triangle scene.  bash tools/gpu_r02m.sh TAG [ncu|noncu] [nobench]
TAG=${1:-r02m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 6 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > $OUT/smoke.log 2>&1; tail -n 2 $OUT/smoke.log
if [ "$3" != "nobench" ]; then
timeout 1500 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 2500 $OUT/bench_default.json; tail -n 3 $OUT/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json
fi
if [ "$2" == "ncu" ]; then
# round 0 has no trace A; -s 3 skips stage1 / trace B / stage2 of round 0 of the first (warm-up) sample: the next 4 matching launches are
# round 1's trace A, stage 1, trace B, stage 2 with every primary-hit path alive
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"wf_trace|wf_stage1|wf_stage2" -s 3 -c 4 -o $OUT/c3s_wavefront -f python bench.py --steps 1 --warmup 1 --spp 1 --no-cpu-baseline --workload c3s > $OUT/ncu_c3s.log 2>&1
tail -n 3 $OUT/ncu_c3s.log | cut -c1-300
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active --clock-control none -c 3000 --csv --log-file $OUT/launches_c4s.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/launches_c4s.log 2>&1
wc -l $OUT/launches_c4s.csv
fi
rm -rf gpurun_out/scenes
ls -la $OUT
