#!/bin/bash
# closing evidence + sanitizer pass on small cases
mkdir -p gpurun_out
timeout -k 5 1100 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v CUDAEvent | tail -3 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout -k 5 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-260 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout -k 5 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 400 -k "(chain16_rdb and 40x40 and conv5_split_x0_only) or (fused_flowstep and main_c12_ragged and split) or shift_first3" 2>&1 | grep -v CUDAEvent | tail -8 | cut -c1-300 | tee gpurun_out/sanitizer_memcheck.log
timeout -k 5 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 400 -k "(chain16_rdb and partial_tiles and conv5_split_x0_only)" 2>&1 | grep -v CUDAEvent | tail -8 | cut -c1-300 | tee gpurun_out/sanitizer_racecheck.log
