#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "config1 or chain16 or native_library or weight_stationary" 2>&1 | grep -v CUDAEvent | tail -2
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
