#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python bench_configs.py --precision tf32x3 2>&1 | grep '^{' | tee gpurun_out/bench_configs_tf32x3.log
timeout -k 5 600 python bench_configs.py --precision tf32 2>&1 | grep '^{' | tee gpurun_out/bench_configs_tf32.log
timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
