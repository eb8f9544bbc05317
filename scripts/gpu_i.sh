#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_pmdi.py -x -q -m gpu -s 2>&1 | tail -15
