#!/bin/bash
mkdir -p gpurun_out /tmp/hcf_script
timeout -k 5 600 python tests/ref_script_worker.py $PWD /tmp/hcf_script 2>&1 | grep -v "CUDAEvent" | tail -30 | cut -c1-260 | tee gpurun_out/ref_script_worker.log
