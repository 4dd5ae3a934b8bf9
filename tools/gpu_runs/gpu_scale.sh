#!/bin/bash
# scaling run of the multi-GPU bench on one box: tools/gpu_scale.sh TAG "2 4 8" [steps]
OUT=gpurun_out/${1:-scale}
mkdir -p $OUT
for N in ${2:-8}; do
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${3:-3} --warmup 3 > $OUT/bench_c5_n$N.json 2> $OUT/bench_c5_n$N.err
python - $OUT/bench_c5_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N",d["n_gpus"],"value",round(d["value"],2),"e2e",round(d["e2e"]["value"],2),"single",d["single_gpu_same_workload"],"identical",d["multi_gpu_bit_identical"],"pass_ms",{k:round(v,1) for k,v in d["pass_ms_per_step"].items()}, d["impl_detail"]["partition"])
except Exception as e:
    print("FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
rm -rf gpurun_out/scenes
