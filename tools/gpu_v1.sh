#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python tools/launch_times.py f16x3 2>&1 | grep -v CUDAEvent | tail -1 | cut -c1-1000 | tee gpurun_out/v1_times.log
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 90 -k "chain16 or config1 or tcgen05 or chained_launch or tensor_core_modes or fcn_as_one or stress_fixture_reverse or weight_stationary" 2>&1 | grep -v CUDAEvent | tail -4 | cut -c1-600 | tee gpurun_out/pytest_v1.log
HCFLOW_LIB=$PWD/hcflow_b200/prof/libhcflow_b200_prof.so HCF_TC_PROF=1 timeout -k 5 200 python tools/prof_chain.py f16x3 2>&1 | grep "chain layers" | cut -c1-900 | tee gpurun_out/v1_prof.log
