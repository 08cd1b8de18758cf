#!/bin/bash
# A/B of the CTA-pair (cta_group::2) conv GEMM: per-op tests, parity tests, then the prod step with and without.
mkdir -p gpurun_out
PHT_OPTIONS=cta_pairs=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv_gemm or padded_conv or padfold or encoder or decoder_tail" > gpurun_out/r2k_ops.log 2>&1; rc=$?; echo "pairs op tests rc=$rc"; tail -5 gpurun_out/r2k_ops.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert" gpurun_out/r2k_ops.log | head -20; exit 1; fi
PHT_OPTIONS=cta_pairs=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_model_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2k_parity.log 2>&1; echo "pairs parity rc=$?"; tail -5 gpurun_out/r2k_parity.log
for opt in "" "cta_pairs=1" "" "cta_pairs=1"; do
  PHT_OPTIONS=$opt timeout 600 python bench.py --no-stock --no-cpu-baseline --no-inference --no-sustained --no-gan-extra > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; python - "$opt" <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r2k_bench.json') if l.startswith('{')][-1])
print('opt=%r value %.1f ms %.3f e2e %.1f' % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value']))
PY
done
PHT_OPTIONS=cta_pairs=1 python tools/d_kernels.py > gpurun_out/r2k_d_kernels_pairs.txt 2>&1; grep -v Warn gpurun_out/r2k_d_kernels_pairs.txt | head -16
python tools/d_kernels.py > gpurun_out/r2k_d_kernels_base.txt 2>&1; grep -v Warn gpurun_out/r2k_d_kernels_base.txt | head -16
