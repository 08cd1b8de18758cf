#!/bin/bash
mkdir -p gpurun_out
for w in dev prod; do for r in 1 0 1 0; do
PHT_FUSED_RING=$r timeout 600 python bench.py --workload $w --no-stock --no-cpu-baseline --no-inference --no-sustained --no-gan-extra > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; python - $w $r <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r2t_bench.json') if l.startswith('{')][-1])
print(sys.argv[1],'ring=%s value %.1f ms %.3f e2e %.1f' % (sys.argv[2], d['value'], d['ms_per_step'], d['e2e']['value']))
PY
done; done
