#!/bin/bash
mkdir -p gpurun_out
( timeout -k 5 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "fused_flowstep and main_c12_ragged and split" 2>&1 | grep -v CUDAEvent | tail -8 | cut -c1-300
  HCF_TC_WS=1 timeout -k 5 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "chain16_rdb and 40x40 and all_split" 2>&1 | grep -v CUDAEvent | tail -8 | cut -c1-300
  HCF_TC_WS=1 timeout -k 5 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "chain16_rdb and 40x40" 2>&1 | grep -v CUDAEvent | tail -5 | cut -c1-300
  timeout -k 5 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "(chain16_rdb and 40x40 and all_split) or (fused_flowstep and main_c12_ragged and inverse-split)" 2>&1 | grep -v CUDAEvent | tail -5 | cut -c1-300 ) | tee gpurun_out/sanitizer_more.log
