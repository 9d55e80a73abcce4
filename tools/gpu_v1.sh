#!/bin/bash
# A/B of a per-item-kernel change through HCF_TC_DEBUG bits: per-launch times
mkdir -p gpurun_out
for d in 0 1024 0 1024; do HCF_TC_DEBUG=$d timeout -k 5 120 python tools/launch_times.py f16x3 2>&1 | grep -v CUDAEvent | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['debug'], {k[:34]:v for k,v in d['ms'].items() if 'chain16' in k or k=='total'})"; done | tee gpurun_out/v1_times.log
