#!/bin/bash
mkdir -p gpurun_out
export PMDI_WATCHDOG_S=600
timeout 120 python scripts/sanitize_small.py spec pool > gpurun_out/k_plain.log 2>&1 && timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py spec pool > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_memcheck.log
