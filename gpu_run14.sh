#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision tf32 --no-modes 2>&1 | tail -1 | tee gpurun_out/bench_tf32.log
