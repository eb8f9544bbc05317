#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_registry.py -x -q -m gpu 2>&1 | tail -25 | cut -c1-250
