#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2n_tests.log
for w in stag dev; do
  timeout 300 python tools/profile_step.py --workload $w > gpurun_out/r2n_plain_$w.log 2>&1; echo "plain $w rc=$?"
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_$w.csv python tools/profile_step.py --workload $w > gpurun_out/r2n_ncu_$w.log 2>&1
  python tools/summarize_launches.py gpurun_out/r2n_launches_$w.csv > gpurun_out/r2n_launch_summary_$w.txt 2>&1; head -14 gpurun_out/r2n_launch_summary_$w.txt; tail -1 gpurun_out/r2n_launch_summary_$w.txt
  timeout 600 python bench.py --workload $w --no-stock --no-cpu-baseline --no-inference --no-sustained --no-gan-extra > gpurun_out/r2n_bench_$w.json 2> gpurun_out/r2n_bench_$w.err; python - $w <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r2n_bench_{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1],'value %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
done
timeout 600 python bench.py --no-stock --no-cpu-baseline --no-inference --no-sustained --no-gan-extra > gpurun_out/r2n_bench_prod.json 2> gpurun_out/r2n_bench_prod.err; python - prod <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r2n_bench_{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1],'value %.1f ms %.3f e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
PY
