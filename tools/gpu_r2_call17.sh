#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "ring or conv_gemm or padded_conv or padfold" > gpurun_out/r2s_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/r2s_ops.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py tests/test_gan_gpu.py -m gpu -q --maxfail=5 > gpurun_out/r2s_model.log 2>&1; echo "model rc=$?"; tail -3 gpurun_out/r2s_model.log
for w in prod dev stag; do
timeout 600 python bench.py --workload $w --no-stock --no-cpu-baseline --no-inference --no-sustained --no-gan-extra > gpurun_out/r2s_bench_$w.json 2> gpurun_out/r2s_bench_$w.err; python - $w <<'PY'
import json,sys
d=json.loads([l for l in open(f'gpurun_out/r2s_bench_{sys.argv[1]}.json') if l.startswith('{')][-1])
print(sys.argv[1],'value %.1f ms %.3f e2e %.1f launches/step %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']//d['steps']))
PY
done
