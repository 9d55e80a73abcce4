#!/bin/bash
# timing experiments on the chained conv kernel (results are WRONG under HCF_TC_DEBUG; timing only)
mkdir -p gpurun_out
run() { env HCF_TC_DEBUG=$1 timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-modes --skip-e2e --precision $2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(json.dumps({'debug':$1,'precision':'$2','ms_per_step':round(d['ms_per_step'],3)}))"; }
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k 'chained or tcgen05 or tensor_core' 2>&1 | tail -3 | tee gpurun_out/pytest_tc.log
{
for dbg in 0 6 14 64 70; do run $dbg tf32; run $dbg tf32x3; done
} | tee gpurun_out/debug_sweep.jsonl
