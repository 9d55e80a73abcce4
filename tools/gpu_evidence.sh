#!/bin/bash
# round-end evidence: full GPU test-suite, smoke, bench (all modes), other configs, ncu launch list + full capture
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout -k 5 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_default.json
cut -c1-300 gpurun_out/bench_default.json
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --precision f16 --no-modes --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_f16.json
timeout -k 5 600 python bench_configs.py --precision f16x3 2>&1 | grep '^{' | tee gpurun_out/bench_configs_f16x3.jsonl
timeout -k 5 600 python bench_configs.py --precision f16 2>&1 | grep '^{' | tee gpurun_out/bench_configs_f16.jsonl
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 25 --csv --log-file gpurun_out/launches_f16x3.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e > /dev/null 2>&1
timeout -k 5 800 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/prof_f16x3_chainL0_v2 python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e 2>&1 | tail -2
ls -la gpurun_out | tail -12
