#!/bin/bash
# bring-up of the fp16 chain kernel: conv-level tests first (own timeout: a hung kernel must not cost the box),
# then end-to-end parity, profile and bench
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chain16_rdb" 2>&1 | tail -15 | tee gpurun_out/pytest_chain16.log
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tensor_core_modes or chained or chain16_full or tcgen05" 2>&1 | tail -15 | tee gpurun_out/pytest_tc.log
for prec in f16 f16x3 tf32 tf32x3; do timeout 120 python tools/prof_chain.py $prec 2>&1 | grep -E "hcf prof|precision"; done | tee gpurun_out/prof_chain.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision f16x3 2>&1 | tail -1 > gpurun_out/bench_f16x3.json
cut -c1-300 gpurun_out/bench_f16x3.json
python -c "
import json
d=json.load(open('gpurun_out/bench_f16x3.json'))
print(d['modes']); print(d['roofline']['conv_by_layer'])"
