#!/bin/bash
# round-2 closing evidence on the final build: full GPU suite, smoke, bench line, ncu --set full of the opt-in
# weight-stationary launch (for the L2 -> SM traffic / tensor-pipe comparison with the per-item schedule)
mkdir -p gpurun_out
timeout -k 5 1100 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v CUDAEvent | tail -4 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout -k 5 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-300 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
HCF_TC_WS=1 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:conv_ws_kernel -s 2 -c 1 -o gpurun_out/prof_ws_chain python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e --no-cpu-baseline 2>&1 | tail -1
ls gpurun_out | tail -8
