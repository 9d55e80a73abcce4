#!/bin/bash
mkdir -p gpurun_out
probe() {  # case passes mt
  HCF_TC_MT=$3 timeout -k 3 30 python tests/tc_probe.py --case $1 --passes $2 2>&1 | grep '^{' | sed "s/^{/{\"mt\": $3, /" || echo "{\"case\": \"$1\", \"passes\": $2, \"mt\": $3, \"hang_or_fail\": true}"
}
{
for c in tiny rdb1 rdb5; do probe $c 1 1; done
for c in tiny rdb1 rdb3 rdb5 prior42; do probe $c 1 2; done
for c in tiny rdb1 rdb5 prior42; do probe $c 3 1; done
} | tee gpurun_out/tc_probe_v2.log
if grep -q hang_or_fail gpurun_out/tc_probe_v2.log; then echo "PROBE FAILED - skipping the rest"; exit 1; fi
timeout -k 5 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision tf32x3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tf32x3.log
