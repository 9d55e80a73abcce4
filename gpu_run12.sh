#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout -k 5 120 python tests/tc_bench.py --precision tf32x3 2>&1 | grep '^{' | tee gpurun_out/tc_bench_x3.log
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_default.log
