#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chained or tensor_core_modes" 2>&1 | tail -12 | tee gpurun_out/pytest_chain.log
if grep -q "failed\|Error\|error" gpurun_out/pytest_chain.log; then echo CHAIN_TEST_FAILED; fi
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_default.log
