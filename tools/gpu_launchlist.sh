#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 27 --csv --log-file gpurun_out/launches_f16x3.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e --no-cpu-baseline > /dev/null 2>&1
grep -c conv_tc_kernel gpurun_out/launches_f16x3.csv
