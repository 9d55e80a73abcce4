#!/bin/bash
# ring-geometry sweep of the chained fp16 kernels (HCF_TC_RINGS = "A stages,B slots,taps per slot")
mkdir -p gpurun_out
run() { env HCF_TC_RINGS=$1 timeout -k 5 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-modes --skip-e2e --precision $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2 rings=$1', round(d['ms_per_step'],3))"; }
{
for r in default 2,6,1 2,5,1 2,3,3 3,2,3 3,4,1; do run $r f16x3; done
for r in default 3,3,3 4,2,9 3,6,3 4,4,3; do run $r f16; done
} | tee gpurun_out/rings_sweep_f16.log
