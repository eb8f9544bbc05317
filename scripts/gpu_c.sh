#!/bin/bash
# 2-GPU box: sharded pool-engine parity + a short sharded bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c_smi.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/c_sharded.log 2>&1; echo "sharded rc=$?"
tail -25 gpurun_out/c_sharded.log | cut -c1-400
