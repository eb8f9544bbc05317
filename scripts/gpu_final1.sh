#!/bin/bash
# one GPU: the whole GPU suite, smoke, the bench of both arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_n1_reference.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().split('\n')[-1])
r=json.loads(open('gpurun_out/r02_bench_n1_reference.json').read().strip().split('\n')[-1])
print('ours ms/step', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], 'kernel', l['roofline']['kernel_ms'], 'ref ms/step', r['ms_per_step'], 'ratio e2e', r['ms_per_step']/l['e2e']['ms_per_step'])
print('same config', l['config']==r['config'])
print('cfg4', l['cfg4']['ms_per_sweep'], l['cfg4']['dense_engine']['ms_per_sweep'], l['cfg4']['dense_engine']['roofline']['frac'], l['cfg4']['dense_engine']['roofline'].get('frac_by_measured_dram_traffic'))
print('pmdi', {k:(round(v['mcmc_iters_per_s'],1), round(v['sweep_device_ms_per_iteration'],1)) for k,v in l['pmdi_end_to_end'].items()})
print('cpu', l['cpu_baseline']['ms_per_sweep'], 'roofline frac', l['roofline']['frac'], 'traffic', l['roofline']['traffic'])
PY
