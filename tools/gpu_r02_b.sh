#!/bin/bash
# round 2: fused FlowStep kernel bring-up
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "fused_flowstep_kernel" 2>&1 | tail -40 | tee gpurun_out/pytest_flowstep.log
if grep -q "failed\|error" gpurun_out/pytest_flowstep.log; then exit 1; fi
timeout -k 5 1700 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout -k 5 700 python bench.py --steps 20 --warmup 5 --no-eager-baseline 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-300 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
