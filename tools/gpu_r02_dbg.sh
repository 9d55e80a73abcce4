#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/debug_sweep.jsonl
for d in 0 2 4 6 8 10 14 64; do HCF_TC_DEBUG=$d timeout 120 python tools/launch_times.py f16x3 2>/dev/null | tail -1 >> gpurun_out/debug_sweep.jsonl; done
cat gpurun_out/debug_sweep.jsonl
