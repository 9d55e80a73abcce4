#!/bin/bash
mkdir -p gpurun_out
for ipp in 8 6 4 3 2; do HCF_WS_IPP=$ipp timeout -k 5 90 python tools/launch_times.py f16x3 2>&1 | grep -v CUDAEvent | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print($ipp, {k[:28]:v for k,v in d['ms'].items() if 'chain16' in k})"; done | tee gpurun_out/ws_ipp_sweep.log
