#!/bin/bash
# A/B on one box: layout variants of the library (scripts/ab/make_variants.py), cfg2 and cfg4 sweeps
mkdir -p gpurun_out
for v in ${VARIANTS:-head old}; do
  cp scripts/ab/lib_$v.so particlemdi.jl_b200/libpmdi_cuda.so
  timeout 100 python scripts/time_configs.py cfg2_multiomics 14 > gpurun_out/ab_$v.log 2>&1
  [ -n "$AB_CFG4" ] && timeout 200 python scripts/time_configs.py cfg4_singlecell 4 > gpurun_out/ab4_$v.log 2>&1 || echo > gpurun_out/ab4_$v.log
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
r=[json.loads(l) for l in open(f'gpurun_out/ab_{v}.log') if l.startswith('{"sweep"')]
r4=[json.loads(l) for l in open(f'gpurun_out/ab4_{v}.log') if l.startswith('{"sweep"')]
print(v, 'cfg2 no-resample', [x['kernel_ms'] for x in r[:-1] if x['resamples']==0], 'all', round(sum(x['kernel_ms'] for x in r[4:-1])/len(r[4:-1]),3), 'cfg4', [x['kernel_ms'] for x in r4[:-1]])
PY
done
