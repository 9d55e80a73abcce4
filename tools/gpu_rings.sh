#!/bin/bash
mkdir -p gpurun_out
run() { env $1 timeout -k 5 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-modes --skip-e2e --precision $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2 $1', round(d['ms_per_step'],3))"; }
{
for r in X=0 HCF_TC_DEBUG=64; do run $r tf32x3; run $r tf32; done
} | tee gpurun_out/rings_sweep3.log
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chained or tcgen05" 2>&1 | tail -3
