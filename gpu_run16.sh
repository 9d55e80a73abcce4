#!/bin/bash
mkdir -p gpurun_out
# launch list of ONE eager pass (default precision): skip the first pass (warm-up), capture the second
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 245 -c 245 --csv --log-file gpurun_out/r01_launches_tf32x3.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e > gpurun_out/ncu_launch.log 2>&1
# full capture: L1 chain (first conv_tc launch) and L0 chain (80th conv_tc launch of a pass)
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 79 -c 1 -o gpurun_out/r01_prof_chain_L0_tf32x3 python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e > gpurun_out/ncu_full1.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 79 -c 1 -o gpurun_out/r01_prof_chain_L0_tf32 python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e --precision tf32 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full1.log gpurun_out/ncu_full2.log gpurun_out/ncu_launch.log
