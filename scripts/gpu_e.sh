#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/e_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/e_bench.json').read().strip().split('\n')[-1])
for k in ['value','ms_per_step','mcmc_sweeps_per_s','gpu_launches','parity','clocks']:
    print(k, l.get(k))
print('e2e', l['e2e'])
print('roofline', {k:v for k,v in l['roofline'].items() if k not in ('note',)})
print('detail', l['detail'])
print('cfg4', json.dumps(l.get('cfg4'))[:3000])
print('pmdi', l.get('pmdi_end_to_end'))
print('cpu', l.get('cpu_baseline'))
PY
