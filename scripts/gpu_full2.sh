#!/bin/bash
# ncu --set full of one settled cfg2 sweep of the spec kernel (after the same command has exited 0 without ncu)
mkdir -p gpurun_out
python scripts/time_configs.py cfg2_multiomics 6 > gpurun_out/f2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sweep_spec$ -s 3 -c 1 -o gpurun_out/r02_spec_cfg2_final -f python scripts/time_configs.py cfg2_multiomics 6 > gpurun_out/f2_ncu.log 2>&1
echo "cfg2 full rc=$?"; ls -la gpurun_out/r02_spec_cfg2_final*
