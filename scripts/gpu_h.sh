#!/bin/bash
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullparity.py tests/test_gpu_pmdi.py -x -q -m gpu 2>&1 | tail -2
for c in cfg2_multiomics cfg4_singlecell; do timeout 300 python scripts/time_configs.py $c 6 2>&1 | tail -2 | cut -c1-260; done
