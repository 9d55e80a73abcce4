#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32x3.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 3 -c 1 -o gpurun_out/prof_conv_tc_v2b_conv3 python tests/tc_bench.py --precision tf32 --mt 1 --only L0_conv3 --reps 2 > gpurun_out/ncu_v2.log 2>&1
tail -2 gpurun_out/ncu_v2.log
