#!/bin/bash
# A/B of the static tile ownership for few-tile chains (HCF_TC_STATIC=0/1), same box, interleaved
mkdir -p gpurun_out
run() { env HCF_TC_STATIC=$1 timeout -k 5 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-modes --precision $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
c={k[8:30]:v['ms'] for k,v in d['roofline']['conv_by_layer'].items()}
print('$2 static=$1', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], c)"; }
{
for rep in 1 2; do for st in 0 1; do run $st f16x3; done; done
for st in 0 1; do run $st f16; done
} | tee gpurun_out/static_ab.log
