#!/bin/bash
# round 2, first GPU call: full GPU suite + smoke + bench (default) + reference arm
mkdir -p gpurun_out
timeout -k 5 1700 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout -k 5 700 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-400 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout -k 5 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference_arm.json
cut -c1-300 gpurun_out/bench_reference_arm.json
