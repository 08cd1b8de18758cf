#!/bin/bash
# A/B: fused 3x3 conv_gemm on the two-group "wide" config (tc_cfg=4) vs the default deep+aux config
mkdir -p gpurun_out
PHT_OPTIONS=tc_cfg=4 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "padfold or padded_conv or baseline_shapes" > gpurun_out/r2e_tests_cfg4.log 2>&1; echo "tests cfg4 rc=$?"; tail -3 gpurun_out/r2e_tests_cfg4.log
for cfg in 0 4 0 4; do
  PHT_OPTIONS=tc_cfg=$cfg timeout 600 python bench.py --no-stock --no-cpu-baseline --no-inference --no-gan-extra --no-sustained --steps 20 > gpurun_out/r2e_bench_cfg$cfg.json 2> gpurun_out/r2e_bench_cfg$cfg.err; echo "cfg $cfg rc=$?: $(head -c 150 gpurun_out/r2e_bench_cfg$cfg.json | cut -c40-150)"
done
PHT_OPTIONS=tc_cfg=4 timeout 300 python tools/profile_step.py > gpurun_out/r2e_plain.log 2>&1 &&
PHT_OPTIONS=tc_cfg=4 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2e_step_cfg4.csv python tools/profile_step.py > gpurun_out/r2e_step_ncu.log 2>&1; echo "step ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2e_step_cfg4.csv | head -12
