#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python tests/precision_probe.py x3_conv5 x3_conv14 tf32x3 2>&1 | grep '^{' | tee gpurun_out/precision_probe_mixed2.log
