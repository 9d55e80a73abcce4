#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python tests/ref_train_worker.py $PWD 2>&1 | grep -v "CUDAEvent\|Warning\|warn" | tail -45 | cut -c1-200 | tee gpurun_out/ref_train_worker.log
