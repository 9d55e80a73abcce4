#!/bin/bash
# round 2: fused FlowStep kernel v2 (8 epilogue warps) -- tests, bench, other configs, ncu of the FlowStep launches
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "fused_flowstep_kernel" 2>&1 | tail -5 | tee gpurun_out/pytest_flowstep.log
if grep -q "failed\|error" gpurun_out/pytest_flowstep.log; then exit 1; fi
timeout -k 5 1700 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout -k 5 700 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-modes 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-300 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout -k 5 600 python bench_configs.py --precision f16x3 2>&1 | grep '^{' | tee gpurun_out/bench_configs_f16x3.jsonl
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 30 --csv --log-file gpurun_out/launches_f16x3.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e > /dev/null 2>&1
timeout -k 5 800 ncu --set full --clock-control none --import-source on -k regex:flowstep_kernel -s 4 -c 4 -o gpurun_out/prof_flowstep python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e 2>&1 | tail -2
ls -la gpurun_out | tail -8
