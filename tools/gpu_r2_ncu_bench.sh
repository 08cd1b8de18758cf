#!/bin/bash
# ncu launch list of the bench command itself (contract: the kernel's share of the step must agree with roofline.share_of_step)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-stock --no-cpu-baseline --no-inference --no-sustained --no-gan-extra"
timeout 300 $CMD > gpurun_out/r2v_bench_plain.json 2> gpurun_out/r2v_bench_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2v_bench_launches.csv $CMD > gpurun_out/r2v_bench_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2v_bench_launches.csv > gpurun_out/r2v_bench_launch_summary.txt 2>&1; head -16 gpurun_out/r2v_bench_launch_summary.txt; tail -1 gpurun_out/r2v_bench_launch_summary.txt
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v_bench_plain.json') if l.startswith('{')][-1])
print('plain: value %.1f share_of_step %.3f' % (d['value'], d['roofline']['share_of_step']))
PY
