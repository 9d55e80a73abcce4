#!/bin/bash
mkdir -p gpurun_out
for dbg in 16 32; do
  for sh in L0_conv1 L0_conv4; do
    HCF_TC_DEBUG=$dbg timeout -k 5 60 python tests/tc_bench.py --precision tf32 --mt 1 --only $sh 2>&1 | grep '^{' | sed "s/^{/{\"dbg\": $dbg, /"
  done
done | tee gpurun_out/tc_bench_debug2.log
timeout -k 5 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32x3.log
