#!/bin/bash
# dev / stag presets at N GPUs (tools/gpu_r2_scale_small.sh N)
N=$1
mkdir -p gpurun_out
for w in dev stag; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --workload $w --no-stock --no-cpu-baseline --no-inference --no-sustained > gpurun_out/r2_bench_${w}_n${N}.json 2> gpurun_out/r2_bench_${w}_n${N}.err; echo "$w rc=$?"
  python - $w $N <<'PY'
import json,sys
d=json.loads([l for l in open(f"gpurun_out/r2_bench_{sys.argv[1]}_n{sys.argv[2]}.json") if l.startswith("{")][-1])
print(sys.argv[1], "N=%s value %.1f ms %.3f e2e %.1f full %.1f" % (sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], d["full_step"]["value"]))
PY
done
