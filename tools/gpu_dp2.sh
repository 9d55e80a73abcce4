#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "dataparallel or nccl" 2>&1 | grep -v CUDAEvent | grep -E "^E|Error|error|assert|passed|failed" | head -30 | cut -c1-300 | tee gpurun_out/pytest_dp2.log
cp gpurun_out/parity_report.json gpurun_out/parity_report_dp2.json 2>/dev/null
