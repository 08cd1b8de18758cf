#!/bin/bash
# Final validation of the round: full GPU test-suite, smoke, the default bench line, launch list + ncu --set full captures.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | tail -12
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.log
timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; python - "$tag" <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/{sys.argv[1]}_bench.json') if l.startswith('{')][-1])
fs=d.get('full_step') or {}
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sustained',d['sustained'] and round(d['sustained']['value'],1),'frac',round(d['roofline']['frac'],3),
      'full',fs and round(fs['value'],1),'inference',d.get('inference') and d['inference'].get('value'),'stock',d.get('stock_cudnn') and d['stock_cudnn'].get('value'),'cpu',d.get('cpu_baseline') and d['cpu_baseline'].get('value'))
PY
timeout 300 python tools/profile_step.py > gpurun_out/${tag}_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_summary.txt 2>&1; head -30 gpurun_out/${tag}_launch_summary.txt
bash tools/gpu_ncu_full.sh ${tag}
for k in conv wgrad attnfwd attnbwd; do python tools/ncu_summary.py gpurun_out/${tag}_${k}.ncu-rep > gpurun_out/${tag}_ncu_full_${k}.txt 2>&1; done
ls -la gpurun_out/${tag}_*.ncu-rep
