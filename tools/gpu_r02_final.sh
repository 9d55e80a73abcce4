#!/bin/bash
# round-2 evidence: full GPU suite, smoke, bench (all modes + baselines), reference arm, other configs, ncu launch list,
# ncu full capture of the dominant launch and of the fused FlowStep launches, wait profile
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout -k 5 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-260 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout -k 5 400 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference_arm.json
cut -c1-200 gpurun_out/bench_reference_arm.json
timeout -k 5 600 python bench_configs.py --precision f16x3 2>&1 | grep '^{' | tee gpurun_out/bench_configs_f16x3.jsonl
timeout -k 5 600 python bench_configs.py --precision f16 2>&1 | grep '^{' | tee gpurun_out/bench_configs_f16.jsonl
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 27 --csv --log-file gpurun_out/launches_f16x3.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e > /dev/null 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 2 -o gpurun_out/prof_f16x3_chains python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e 2>&1 | tail -1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:flowstep_kernel -s 4 -c 4 -o gpurun_out/prof_flowstep python bench.py --steps 1 --warmup 1 --no-graph --skip-e2e 2>&1 | tail -1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 2 -o gpurun_out/prof_rescaling_dense python tools/launch_times.py f16x3 rescaling_x4 forward 64 64 2>&1 | tail -1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:flowstep_kernel -c 2 -o gpurun_out/prof_rescaling_flowstep_fwd python tools/launch_times.py f16x3 rescaling_x4 forward 64 64 2>&1 | tail -1
for d in forward reverse; do LT_MIN_MS=0.15 timeout 200 python tools/launch_times.py f16x3 rescaling_x4 $d 64 64 2>/dev/null | tail -1; done > gpurun_out/launch_times_rescaling.jsonl
LT_MIN_MS=0.1 timeout 200 python tools/launch_times.py f16x3 sr_x8 reverse 32 20 2>/dev/null | tail -1 > gpurun_out/launch_times_sr_x8.jsonl
HCFLOW_LIB=$PWD/hcflow_b200/prof/libhcflow_b200_prof.so HCF_TC_PROF=1 timeout -k 5 200 python tools/prof_chain.py f16x3 2> gpurun_out/wait_profile_f16x3.log
ls gpurun_out | tail -14
