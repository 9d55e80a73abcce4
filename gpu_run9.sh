#!/bin/bash
mkdir -p gpurun_out
for dbg in 2 4 6 8 14; do
  for sh in L0_conv1 L0_conv4 L0_conv5 L1_conv4; do
    HCF_TC_DEBUG=$dbg timeout -k 5 60 python tests/tc_bench.py --precision tf32 --mt 1 --only $sh 2>&1 | grep '^{' | sed "s/^{/{\"dbg\": $dbg, /"
  done
done | tee gpurun_out/tc_bench_debug.log
