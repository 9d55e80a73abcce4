#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v CUDAEvent | tail -4 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
