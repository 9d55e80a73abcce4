#!/bin/bash
mkdir -p gpurun_out
# profiling build first (in-kernel wait profiler), then the product build
HCF_BUILD_PROF=1 python -m hcflow_b200.build --force > /dev/null 2>&1
for prec in f16 f16x3; do timeout 120 python tools/prof_chain.py $prec 2>&1 | grep -E "hcf prof|precision"; done | tee gpurun_out/prof_chain3.log
python -m hcflow_b200.build --force > /dev/null 2>&1
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core_modes or chained or chain16 or tcgen05" 2>&1 | tail -4 | tee gpurun_out/pytest_tc.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision f16x3 2>&1 | tail -1 > gpurun_out/bench_f16x3.json
python -c "
import json
d=json.load(open('gpurun_out/bench_f16x3.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['modes'].items()}); print({k[:22]:v['ms'] for k,v in d['roofline']['conv_by_layer'].items()}); print(d['e2e'])"
