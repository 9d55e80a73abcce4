#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout -k 5 600 python bench_configs.py --precision tf32x3 2>&1 | grep '^{' | tee gpurun_out/bench_configs_tf32x3.log
timeout -k 5 600 python bench_configs.py --precision tf32 2>&1 | grep '^{' | tee gpurun_out/bench_configs_tf32.log
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/bench_default.log
