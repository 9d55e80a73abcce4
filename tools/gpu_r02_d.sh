#!/bin/bash
# wait profile of the fused FlowStep launches and the encoder chains (prof build of the library)
mkdir -p gpurun_out
HCFLOW_LIB=$PWD/hcflow_b200/prof/libhcflow_b200_prof.so HCF_TC_PROF=1 timeout -k 5 300 python tools/prof_chain.py f16x3 2> gpurun_out/wait_profile_f16x3.log
cat gpurun_out/wait_profile_f16x3.log
