#!/bin/bash
# closing evidence on the final build: full GPU suite, smoke, bench line, launch list, ncu of the dominant launch, wait profile
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v CUDAEvent | tail -3 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout -k 5 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-260 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 27 --csv --log-file gpurun_out/launches_f16x3.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 2 -o gpurun_out/prof_f16x3_chains python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e --no-cpu-baseline 2>&1 | tail -1
HCFLOW_LIB=$PWD/hcflow_b200/prof/libhcflow_b200_prof.so HCF_TC_PROF=1 timeout -k 5 200 python tools/prof_chain.py f16x3 2> gpurun_out/wait_profile_f16x3.log
timeout -k 5 600 python bench_configs.py --precision f16x3 2>&1 | grep '^{' | tee gpurun_out/bench_configs_f16x3.jsonl
ls gpurun_out | tail -5
