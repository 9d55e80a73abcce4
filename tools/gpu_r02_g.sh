#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "inverse_path" 2>&1 | tail -30 | tee gpurun_out/pytest_tiling.log
