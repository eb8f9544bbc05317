#!/bin/bash
# bench at N = $1 GPUs (torchrun), tight timeout
N=$1
mkdir -p gpurun_out
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
tail -3 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    l=json.loads([x for x in open(f'gpurun_out/r02_bench_n{N}.json').read().strip().split('\n') if x.startswith('{')][-1])
    print('value',l['value'],'ms_per_step',l['ms_per_step'],'e2e_ms',l['e2e']['ms_per_step'],'kernel_ms',l['roofline']['kernel_ms'])
    print('timed',l['ms_per_timed_step'])
    print('parity',l['parity'])
    c=l.get('cfg4_strong'); print('cfg4_strong', c and (c['ms_per_sweep'], c['ms_per_timed_sweep'], c['engine']))
    print('detail', {k:l['detail'][k] for k in ('engine','rows_pulled_from_peers_per_sweep','resamples_per_sweep','distinct_clusters_evaluated_per_sweep')})
except Exception as e:
    print('no line', e)
PY
