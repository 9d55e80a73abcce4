#!/bin/bash
mkdir -p gpurun_out
run() { env HCF_TC_DEBUG=$1 timeout -k 5 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-modes --precision f16x3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
c={k[8:30]:v['ms'] for k,v in d['roofline']['conv_by_layer'].items()}
print('debug=$1', round(d['ms_per_step'],3), c)"; }
{ run 0; run 128; } | tee gpurun_out/stepcost.log
