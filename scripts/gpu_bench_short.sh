#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu --no-pmdi > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/q_bench.json').read().strip().split('\n')[-1])
print('ms/step', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], 'kernel', l['roofline']['kernel_ms'], 'timed', l['ms_per_timed_step'])
print('cfg4', l['cfg4']['ms_per_sweep'], l['cfg4'].get('ms_per_timed_sweep'))
print('phases', l['roofline'].get('warp_ms_mean_over_ctas'))
PY
