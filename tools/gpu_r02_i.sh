#!/bin/bash
# quick check of a FlowStep-kernel change: its parity tests, the stress fixtures, per-launch times, wait profile
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "flowstep or stress_fixture or config1 or tensor_core_modes" 2>&1 | tail -3
timeout 100 python tools/launch_times.py f16x3 2>/dev/null | tail -1 | cut -c1-700
HCFLOW_LIB=$PWD/hcflow_b200/prof/libhcflow_b200_prof.so HCF_TC_PROF=1 timeout -k 5 200 python tools/prof_chain.py f16x3 2>&1 | grep "flowstep chain" | cut -c1-900
