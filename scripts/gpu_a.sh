#!/bin/bash
# round-2 GPU call A: first run of the pool engine
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
tail -5 gpurun_out/a_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/a_parity.log 2>&1; echo "parity rc=$?"
tail -30 gpurun_out/a_parity.log
timeout 600 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_fullsize.py tests/test_gpu_pmdi.py -q -m gpu > gpurun_out/a_rest.log 2>&1; echo "rest rc=$?"
tail -30 gpurun_out/a_rest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
