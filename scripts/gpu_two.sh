#!/bin/bash
# two GPUs: the sharded parity tests, per-sweep sub-phase times, then the bench at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/time_sharded.py cfg2_multiomics 12 P=512 > gpurun_out/two_sharded.log 2>&1
grep '"rank": 0' gpurun_out/two_sharded.log | tail -8 | cut -c1-330
bash scripts/gpu_n.sh 2
