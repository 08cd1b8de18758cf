#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --maxfail=10 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/r2g_pytest.log | tail -12
for i in 1 2; do
timeout 600 python bench.py --no-stock --no-cpu-baseline --no-inference --no-gan-extra --no-sustained --steps 20 > gpurun_out/r2g_bench_$i.json 2> gpurun_out/r2g_bench_$i.err; echo "bench rc=$?: $(head -c 150 gpurun_out/r2g_bench_$i.json | cut -c40-150)"; tail -2 gpurun_out/r2g_bench_$i.err
done
timeout 300 python tools/profile_step.py > gpurun_out/r2g_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2g_step_ncu.csv python tools/profile_step.py > gpurun_out/r2g_step_ncu.log 2>&1; echo "step ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2g_step_ncu.csv > gpurun_out/r2g_launches.txt; head -16 gpurun_out/r2g_launches.txt; tail -1 gpurun_out/r2g_launches.txt
