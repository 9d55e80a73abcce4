#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python tests/precision_probe.py tf32x3 2>&1 | grep '^{' | tee gpurun_out/precision_probe_x3.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32x3.log
# ncu: launch list of one eager step, then a full capture of the conv kernel (3 launches of a mid-size layer)
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 1 --warmup 1 --precision tf32 --no-graph --skip-e2e > gpurun_out/ncu_launch.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 900 -c 3 -o gpurun_out/prof_conv_tc_r01a python bench.py --steps 1 --warmup 1 --precision tf32 --no-graph --skip-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
