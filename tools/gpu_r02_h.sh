#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "chain16_rdb or fcn_as_one_chain" 2>&1 | tail -3
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "chain16_full or tensor_core_modes or config1 or stress_fixture_reverse" 2>&1 | tail -3
timeout 100 python tools/launch_times.py f16x3 2>/dev/null | tail -1 | cut -c1-520
HCF_TC_PAIR=0 timeout 100 python tools/launch_times.py f16x3 2>/dev/null | tail -1 | cut -c1-520
