#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32.log
