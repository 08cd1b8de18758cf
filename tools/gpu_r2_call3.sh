#!/bin/bash
# round 2, GPU call 3: new attention backward (TMA reduce-add, ordered) -- kernel tests first, then everything
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -s -k "attention" > gpurun_out/r2c_attn.log 2>&1; rc=$?; echo "attention tests rc=$rc"; tail -15 gpurun_out/r2c_attn.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|worst gradient|PSNR|capture failed|Error" gpurun_out/r2c_pytest.log | tail -30
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c_smoke.log
timeout 900 python bench.py --no-stock --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; head -c 1200 gpurun_out/r2c_bench.json; echo; tail -5 gpurun_out/r2c_bench.err
PHT_OPTIONS=attn_bwd_direct=0 timeout 900 python bench.py --no-stock --no-cpu-baseline --no-inference --no-gan-extra --no-sustained > gpurun_out/r2c_bench_fold.json 2> gpurun_out/r2c_bench_fold.err; echo "bench fold-mode rc=$?"; head -c 400 gpurun_out/r2c_bench_fold.json; echo; python tools/diag_attn.py > gpurun_out/r2c_diag_attn.txt 2>&1; cat gpurun_out/r2c_diag_attn.txt
timeout 300 python tools/profile_step.py > gpurun_out/r2c_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2c_step_ncu.csv python tools/profile_step.py > gpurun_out/r2c_step_ncu.log 2>&1; echo "step ncu rc=$?"
python tools/summarize_bandwidth_ncu.py gpurun_out/r2c_step_ncu.csv 'pht::' | head -30
