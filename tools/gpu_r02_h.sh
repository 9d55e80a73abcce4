#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "chained_launch_is_bit or chain16_full or tensor_core_modes or config1" 2>&1 | tail -3
timeout 120 python tools/launch_times.py f16x3 2>/dev/null | tail -1 | cut -c1-520
