#!/bin/bash
# two GPUs: the sharded parity tests, then the bench at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
bash scripts/gpu_n.sh 2
