#!/bin/bash
mkdir -p gpurun_out
{
HCF_TC_DEBUG=1 timeout -k 5 120 python tests/tc_bench.py --precision tf32 --mt 1
} 2>&1 | grep '^{' | tee gpurun_out/tc_bench_aligned.log
