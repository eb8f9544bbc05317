#!/bin/bash
# quick check of a kernel change: parity tests, cfg2 / cfg4 sweep times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullparity.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python scripts/time_configs.py cfg2_multiomics 14 > gpurun_out/q_cfg2.log 2>&1; tail -7 gpurun_out/q_cfg2.log | cut -c1-400
timeout 300 python scripts/time_configs.py cfg4_singlecell 5 > gpurun_out/q_cfg4.log 2>&1; tail -3 gpurun_out/q_cfg4.log | cut -c1-400
