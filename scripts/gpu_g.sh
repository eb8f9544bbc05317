#!/bin/bash
# profiles of round 2: launch list of the bench command, ncu --set full of the sweep kernel (cfg2, cfg4)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-cfg4 --no-pmdi --no-parity"
$CMD > gpurun_out/g_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/g_ncu_l.log 2>&1
echo "launch list rc=$?"
python scripts/time_configs.py cfg2_multiomics 6 > gpurun_out/g_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sweep_spec$ -s 3 -c 1 -o gpurun_out/r02_spec_cfg2 -f python scripts/time_configs.py cfg2_multiomics 6 > gpurun_out/g_ncu2.log 2>&1
echo "cfg2 full rc=$?"
python scripts/time_configs.py cfg4_singlecell 4 > gpurun_out/g_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sweep_spec$ -s 2 -c 1 -o gpurun_out/r02_spec_cfg4 -f python scripts/time_configs.py cfg4_singlecell 4 > gpurun_out/g_ncu4.log 2>&1
echo "cfg4 full rc=$?"
ls -la gpurun_out/r02_*
