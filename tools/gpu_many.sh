#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "sample_many" 2>&1 | grep -v CUDAEvent | grep -E "^E|Error|error|assert|passed|failed" | head -12 | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/parity_report.json')); print(d.get('sample_many'))"
