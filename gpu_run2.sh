#!/bin/bash
mkdir -p gpurun_out
for c in tiny rdb1 rdb3 rdb5 prior42 fcn3_22; do
  timeout -k 3 25 python tests/tc_probe.py --case $c --passes 3 2>&1 | grep '^{' || echo "{\"case\": \"$c\", \"passes\": 3, \"hang_or_fail\": true}"
done | tee gpurun_out/tc_probe3.log
timeout -k 5 300 python tests/precision_probe.py fp32 tf32 2>&1 | grep '^{' | tee gpurun_out/precision_probe.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_tf32.log
