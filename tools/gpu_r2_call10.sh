#!/bin/bash
mkdir -p gpurun_out
for ks in 1 3; do timeout 120 python tools/diag_chain.py --ks $ks 2>&1 | grep -v Warn; done
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv_gemm or padded_conv or padfold or encoder or decoder_tail" > gpurun_out/r2l_ops.log 2>&1; rc=$?; echo "op tests rc=$rc"; tail -2 gpurun_out/r2l_ops.log
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_model_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2l_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/r2l_parity.log
for opt in "" "cta_pairs=1" ""; do
  PHT_OPTIONS=$opt timeout 600 python bench.py --no-stock --no-cpu-baseline --no-inference --no-sustained --no-gan-extra > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; python - "$opt" <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r2l_bench.json') if l.startswith('{')][-1])
print('opt=%r value %.1f ms %.3f e2e %.1f' % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value']))
PY
done
