#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "bn_act or critic or colsum" > gpurun_out/r2q_bn.log 2>&1; echo "bn tests rc=$?"; tail -2 gpurun_out/r2q_bn.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "gan or critic or Gan or GAN" > gpurun_out/r2q_gan.log 2>&1; echo "gan tests rc=$?"; tail -2 gpurun_out/r2q_gan.log
python tools/d_kernels.py > gpurun_out/r2q_d_kernels.txt 2>&1; grep -v Warn gpurun_out/r2q_d_kernels.txt | head -8; grep -E "Memset|Memcpy" gpurun_out/r2q_d_kernels.txt
timeout 900 python bench.py --no-stock --no-cpu-baseline --no-inference --no-sustained > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2q_bench.json') if l.startswith('{')][-1])
fs=d['full_step']
print('value %.1f full %.1f (%.2f ms) full e2e %.1f' % (d['value'], fs['value'], fs['ms_per_step'], fs['e2e']['value']))
PY
