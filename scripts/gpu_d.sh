#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullparity.py -x -q -m gpu 2>&1 | tail -3
for c in cfg2_multiomics cfg3_tcga; do timeout 600 python scripts/time_configs.py $c 6 2>&1 | tail -3 | cut -c1-330; done
PMDI_ENGINE=pool timeout 600 python scripts/time_configs.py cfg3_tcga 6 2>&1 | tail -2 | cut -c1-330
PMDI_ENGINE=pool timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullparity.py -x -q -m gpu 2>&1 | tail -3
PMDI_ENGINE=dense timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
