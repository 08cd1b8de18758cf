#!/bin/bash
mkdir -p gpurun_out
for o in "" half_ring=1; do echo "== opt=$o"; PHT_OPTIONS=$o timeout 120 python tools/diag_conv.py --ks 3 2>&1 | grep -v Warn | grep -E "median|tile [1-3]:"; done
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "pipeline_variants" > gpurun_out/r2o_var.log 2>&1; echo "variants rc=$?"; tail -3 gpurun_out/r2o_var.log
PHT_OPTIONS=half_ring=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv_gemm or padded_conv or padfold or encoder or decoder_tail" > gpurun_out/r2o_ops.log 2>&1; rc=$?; echo "half_ring op tests rc=$rc"; tail -2 gpurun_out/r2o_ops.log
for opt in "" "half_ring=1" "" "half_ring=1"; do
  PHT_OPTIONS=$opt timeout 600 python bench.py --no-stock --no-cpu-baseline --no-inference --no-sustained --no-gan-extra > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; python - "$opt" <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r2o_bench.json') if l.startswith('{')][-1])
print('opt=%r value %.1f ms %.3f e2e %.1f frac %.3f' % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
PY
done
