#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "large_ragged" 2>&1 | grep -v CUDAEvent | grep -E "^E|Error|error|assert|passed|failed" | head -12 | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/parity_report.json')); print(d.get('large_ragged_250x182'))"
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
