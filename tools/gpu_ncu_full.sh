#!/bin/bash
# ncu --set full captures of the hot kernels inside one training step (tools/profile_step.py); summaries by
# tools/ncu_summary.py / tools/ncu_stalls.py.   tools/gpu_ncu_full.sh <tag>
tag=$1
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
timeout 300 $NCU -k regex:conv_gemm_tc --launch-skip 5 --launch-count 5 -o gpurun_out/${tag}_conv python tools/profile_step.py > gpurun_out/${tag}_ncu_conv.log 2>&1; echo "conv rc=$?"
timeout 300 $NCU -k regex:wgrad_tc --launch-skip 1 --launch-count 1 -o gpurun_out/${tag}_wgrad python tools/profile_step.py > gpurun_out/${tag}_ncu_wgrad.log 2>&1; echo "wgrad rc=$?"
timeout 300 $NCU -k regex:attn_fwd_tc -c 1 -o gpurun_out/${tag}_attnfwd python tools/profile_step.py > gpurun_out/${tag}_ncu_attnfwd.log 2>&1; echo "attnfwd rc=$?"
timeout 300 $NCU -k regex:attn_bwd_tc -c 1 -o gpurun_out/${tag}_attnbwd python tools/profile_step.py > gpurun_out/${tag}_ncu_attnbwd.log 2>&1; echo "attnbwd rc=$?"
