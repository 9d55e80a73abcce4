#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v CUDAEvent | tail -3 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
timeout -k 5 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout -k 5 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-220 gpurun_out/bench_default.json
