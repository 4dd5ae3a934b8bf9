#!/bin/bash
# ncu scheduler / warp-state / instruction-cache counters of small vs large kernels (instruction-cache hypothesis).
TAG=${1:-ncu_small}
OUT=gpurun_out/$TAG
mkdir -p $OUT
M="smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,sm__icc_request_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread"
for w in c1 c2 c4; do
  timeout 600 ncu --metrics $M --clock-control none -s 20 -c 12 --csv --log-file $OUT/metrics_$w.csv python bench.py --steps 1 --warmup 3 --spp 1 --no-cpu-baseline --workload $w > $OUT/log_$w.txt 2>&1
done
python - <<'PY'
import csv,glob,collections
for f in sorted(glob.glob("gpurun_out/*/metrics_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if not rows: continue
    h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
    d=collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((r[ii], r[ki][:90]),{})[r[mi]]=r[vi]
    seen=set()
    for (i,k),m in d.items():
        if k in seen: continue
        seen.add(k)
        print(f, k)
        print("   ", {a.split('.')[0].replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active',''): b for a,b in m.items()})
PY
