#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python tests/precision_probe.py enc1_fcn3 enc3_fcn1 2>&1 | grep '^{' | tee gpurun_out/precision_probe_mixed.log
