#!/bin/bash
# 2-GPU box: sharded parity tests, then the bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/f_bench2.json 2> gpurun_out/f_bench2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/f_bench2.err | cut -c1-300
python - <<'PY'
import json
try:
    l=json.loads([x for x in open('gpurun_out/f_bench2.json').read().strip().split('\n') if x.startswith('{')][-1])
    for k in ['value','ms_per_step','n_gpus','parity']:
        print(k, l.get(k))
    print('e2e', l['e2e'])
    print('detail', l['detail'])
    print('cfg4_strong', json.dumps(l.get('cfg4_strong'))[:1500])
except Exception as e:
    print('no line', e)
PY
