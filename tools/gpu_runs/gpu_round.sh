#!/bin/bash
# One GPU-box evidence session: parity tests, the driver's two bench arms, one line per workload, ncu launch list + full captures.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 6 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > $OUT/smoke.log 2>&1; tail -n 2 $OUT/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
for w in c1 c3 c4; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
for w in c3s c4s; do
  timeout 900 python bench.py --steps 2 --warmup 3 --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
for form in persistent tiles; do
  B2PT_MEGAKERNEL=$form timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_$form.json 2> $OUT/bench_c2_$form.err
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"],2), d["unit"], "e2e", round(d["e2e"]["value"],2), "mk_ms", d.get("pass_ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
# ncu: launch list of the bench command, then full captures of the dominant kernel on both scenes
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c2.csv python bench.py --steps 2 --warmup 3 --spp 2 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_phased -s 3 -c 1 -o $OUT/c2_megakernel -f python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline > $OUT/ncu_full_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel_phased -s 3 -c 1 -o $OUT/c3s_megakernel -f python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload c3s > $OUT/ncu_full_c3s.log 2>&1
rm -rf gpurun_out/scenes
ls -la $OUT | head -40
