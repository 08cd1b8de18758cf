#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/r2j_pytest.log | tail -12
python tools/d_kernels.py > gpurun_out/r2j_d_kernels.txt 2>&1; head -14 gpurun_out/r2j_d_kernels.txt | grep -v Warn
timeout 900 python bench.py --no-stock --no-cpu-baseline --no-inference > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2j_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench.json') if l.startswith('{')][-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sustained',round(d['sustained']['value'],1),'full',round(d['full_step']['value'],1),round(d['full_step']['ms_per_step'],2),'ms; full e2e',round(d['full_step']['e2e']['value'],1))
PY
PHT_CRITIC_FUSED_BN=0 timeout 900 python bench.py --no-stock --no-cpu-baseline --no-inference --no-sustained > gpurun_out/r2j_bench_unfused.json 2> gpurun_out/r2j_bench_unfused.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench_unfused.json') if l.startswith('{')][-1])
print('UNFUSED BN: value',round(d['value'],1),'full',round(d['full_step']['value'],1),round(d['full_step']['ms_per_step'],2),'ms')
PY
