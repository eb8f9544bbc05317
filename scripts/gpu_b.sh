#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/b_parity.log 2>&1; echo "parity rc=$?"
tail -15 gpurun_out/b_parity.log
timeout 1200 python -m pytest tests/test_gpu_fullparity.py -q -m gpu -s > gpurun_out/b_full.log 2>&1; echo "fullparity rc=$?"
grep -v "^$" gpurun_out/b_full.log | tail -40 | cut -c1-600
timeout 300 python scripts/time_configs.py cfg2_multiomics 10 > gpurun_out/b_t_cfg2.log 2>&1; tail -4 gpurun_out/b_t_cfg2.log
timeout 600 python scripts/time_configs.py cfg4_singlecell 6 > gpurun_out/b_t_cfg4.log 2>&1; tail -8 gpurun_out/b_t_cfg4.log
timeout 300 python scripts/time_configs.py cfg3_tcga 5 > gpurun_out/b_t_cfg3.log 2>&1; tail -3 gpurun_out/b_t_cfg3.log
timeout 300 python scripts/time_configs.py cfg5_scaling 5 P=4096 > gpurun_out/b_t_cfg5.log 2>&1; tail -3 gpurun_out/b_t_cfg5.log
