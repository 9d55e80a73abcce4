#!/bin/bash
mkdir -p gpurun_out
{
timeout -k 5 120 python tests/tc_bench.py --precision tf32 --mt 1
timeout -k 5 120 python tests/tc_bench.py --precision tf32 --mt 2
timeout -k 5 120 python tests/tc_bench.py --precision tf32 --mt 1 --flush
timeout -k 5 120 python tests/tc_bench.py --precision tf32x3
} 2>&1 | grep '^{' | tee gpurun_out/tc_bench.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 3 -c 1 -o gpurun_out/prof_conv_tc_v2_conv3 python tests/tc_bench.py --precision tf32 --mt 1 --only L0_conv3 --reps 2 > gpurun_out/ncu_v2.log 2>&1
tail -3 gpurun_out/ncu_v2.log
