#!/bin/bash
# One GPU-box round: GPU tests, a plain bench run, then the ncu launch list of one training step.
#   tools/gpu_round.sh <tag> [pytest -k expression]
tag=$1; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/${tag}_pytest.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
fi
echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-gan-extra > gpurun_out/${tag}_bench_prod.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/${tag}_bench_prod.log | cut -c1-400
timeout 300 python tools/profile_step.py > gpurun_out/${tag}_plain.log 2>&1
echo "plain rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_summary.txt 2>&1
head -24 gpurun_out/${tag}_launch_summary.txt
