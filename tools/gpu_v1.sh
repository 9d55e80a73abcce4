#!/bin/bash
# per-launch times of configs[1] and the other BASELINE configs; chain parity tests
mkdir -p gpurun_out
timeout -k 5 120 python tools/launch_times.py f16x3 2>&1 | grep -v CUDAEvent | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k[:34]:v for k,v in d['ms'].items() if 'chain16' in k or k=='total'})" | tee gpurun_out/v1_times.log
timeout -k 5 400 python bench_configs.py --precision f16x3 2>&1 | grep '^{' | tee gpurun_out/bench_configs_f16x3.jsonl
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "chain16 or config or tensor_core_modes or stress_fixture_reverse or ragged or range_guard" 2>&1 | grep -v CUDAEvent | tail -2 | cut -c1-300
