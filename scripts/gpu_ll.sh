#!/bin/bash
# launch list of the bench command (after the same command has exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-cfg4 --no-pmdi --no-parity"
$CMD > gpurun_out/ll_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ll_ncu.log 2>&1
echo "launch list rc=$?"; tail -c 400 gpurun_out/ll_plain.log
