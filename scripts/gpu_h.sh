#!/bin/bash
# 1 GPU: parity of the default engines and a short bench (after the counters-run-on change)
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullparity.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cfg4 --no-pmdi > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/h_bench.err
python -c "
import json
l=json.loads(open('gpurun_out/h_bench.json').read().strip().split('\n')[-1])
print(l['ms_per_step'], l['e2e']['ms_per_step'], l['parity']['ok'], l['roofline']['kernel_ms'])"
