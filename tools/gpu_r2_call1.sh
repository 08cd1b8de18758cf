#!/bin/bash
# round 2, GPU call 1: full GPU test suite (all failures), default bench, dev/stag bench, bandwidth kernels (plain + ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|worst gradient|PSNR|fp32 \[|bf16 \[|^rank" gpurun_out/r2a_pytest.log | tail -60
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a_smoke.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
for w in dev stag; do
  timeout 600 python bench.py --workload $w --no-inference --no-stock --no-cpu-baseline > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err; echo "bench $w rc=$?"; cut -c1-900 gpurun_out/r2a_bench_$w.json
done
PHT_BW_ITERS=2 timeout 300 python tools/bench_bandwidth_kernels.py > gpurun_out/r2a_bw_plain.txt 2>&1 &&
PHT_BW_ITERS=2 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2a_bw_ncu.csv python tools/bench_bandwidth_kernels.py > gpurun_out/r2a_bw_ncu.log 2>&1; echo "bw ncu rc=$?"
timeout 300 python tools/profile_step.py > gpurun_out/r2a_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2a_step_ncu.csv python tools/profile_step.py > gpurun_out/r2a_step_ncu.log 2>&1; echo "step ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2a_step_ncu.csv | head -30
