#!/bin/bash
# 2 GPUs: the whole GPU suite (1-GPU tests + 2-rank sharded ones)
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
