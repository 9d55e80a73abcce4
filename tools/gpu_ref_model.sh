#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "reference_model_wrapper or dataparallel" 2>&1 | grep -v CUDAEvent | grep -E "^E|Error|error|assert|passed|failed|OK" | head -30 | cut -c1-400 | tee gpurun_out/pytest_ref_model.log
