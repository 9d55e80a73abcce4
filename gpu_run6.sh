#!/bin/bash
mkdir -p gpurun_out
probe() {  # case passes mt
  HCF_TC_MT=$3 timeout -k 3 30 python tests/tc_probe.py --case $1 --passes $2 2>&1 | grep '^{' | sed "s/^{/{\"mt\": $3, /" || echo "{\"case\": \"$1\", \"passes\": $2, \"mt\": $3, \"hang_or_fail\": true}"
}
{
for c in rdb1 rdb5; do probe $c 1 1; done
for c in rdb3 prior42; do probe $c 1 2; done
for c in rdb1 prior42; do probe $c 3 1; done
} | tee gpurun_out/tc_probe_v2.log
if grep -q hang_or_fail gpurun_out/tc_probe_v2.log; then echo "PROBE FAILED - skipping the rest"; exit 1; fi
{
timeout -k 5 120 python tests/tc_bench.py --precision tf32 --mt 1
timeout -k 5 120 python tests/tc_bench.py --precision tf32 --mt 2
timeout -k 5 120 python tests/tc_bench.py --precision tf32x3
} 2>&1 | grep '^{' | tee gpurun_out/tc_bench.log
