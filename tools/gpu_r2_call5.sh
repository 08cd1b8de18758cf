#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/r2f_pytest.log | tail -12
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"
for i in 1 2; do
timeout 600 python bench.py --no-stock --no-cpu-baseline --no-inference --no-gan-extra --no-sustained --steps 20 > gpurun_out/r2f_bench_$i.json 2> gpurun_out/r2f_bench_$i.err; echo "bench rc=$?: $(head -c 150 gpurun_out/r2f_bench_$i.json | cut -c40-150)"; tail -2 gpurun_out/r2f_bench_$i.err
done
timeout 300 python tools/profile_step.py > gpurun_out/r2f_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2f_step_ncu.csv python tools/profile_step.py > gpurun_out/r2f_step_ncu.log 2>&1; echo "step ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2f_step_ncu.csv > gpurun_out/r2f_launches.txt; head -16 gpurun_out/r2f_launches.txt; tail -1 gpurun_out/r2f_launches.txt
