#!/bin/bash
# round 2, GPU call 2 (2 GPUs): GPU test suite incl. the NCCL DP parity, smoke, bench at N=1 and N=2 with graph replay
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|worst gradient|PSNR|^rank|capture failed|Error" gpurun_out/r2b_pytest.log | tail -40
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2b_smoke.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; head -c 1800 gpurun_out/r2b_bench.json; echo; grep -o '"full_step".\{0,400\}' gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
PHT_STEP_GRAPH=0 timeout 900 python bench.py --no-inference --no-stock --no-cpu-baseline --no-gan-extra > gpurun_out/r2b_bench_eager.json 2> gpurun_out/r2b_bench_eager.err; echo "bench eager rc=$?"; head -c 400 gpurun_out/r2b_bench_eager.json; echo
for w in dev stag; do
  timeout 600 python bench.py --workload $w --no-inference --no-stock --no-cpu-baseline > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err; echo "bench $w rc=$?"; head -c 700 gpurun_out/r2b_bench_$w.json; echo;  grep -o '"full_step".\{0,200\}' gpurun_out/r2b_bench_$w.json; tail -3 gpurun_out/r2b_bench_$w.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --no-stock > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench n2 rc=$?"; head -c 1500 gpurun_out/r2b_bench_n2.json; echo; grep -o '"full_step".\{0,300\}' gpurun_out/r2b_bench_n2.json; tail -5 gpurun_out/r2b_bench_n2.err
