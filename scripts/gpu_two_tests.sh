#!/bin/bash
# two GPUs: the sharded parity tests only
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
