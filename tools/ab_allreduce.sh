#!/bin/bash
# A/B of the gradient all-reduce schedule at N GPUs (default 2): overlapped buckets vs one all-reduce after backward.
N=${1:-2}
mkdir -p gpurun_out
for mode in overlap end; do
  PHT_GRAD_ALLREDUCE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29700 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 4 --no-cpu-baseline --no-gan-extra --no-inference \
    > gpurun_out/ab_allreduce_${mode}_n$N.json 2> gpurun_out/ab_allreduce_${mode}_n$N.err
  echo "$mode rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/ab_allreduce_${mode}_n$N.json").read().strip().splitlines()[-1])
    print("$mode", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["clocks"])
except Exception as e:
    print("parse failed", e)
P
  tail -2 gpurun_out/ab_allreduce_${mode}_n$N.err
done
