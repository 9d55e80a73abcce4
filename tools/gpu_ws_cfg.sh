#!/bin/bash
# the weight-stationary schedule on the other BASELINE configs (many image groups per layer: x8 B=32, rescaling B=64)
mkdir -p gpurun_out
for ws in 0 1; do HCF_TC_WS=$ws timeout -k 5 400 python bench_configs.py --precision f16x3 2>&1 | grep '^{' | sed "s/^{/{\"ws\": $ws, /"; done | tee gpurun_out/ws_other_configs.jsonl
