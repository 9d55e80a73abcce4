#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core_modes or chained or chain16 or fcn_as_one or reverse_matches" 2>&1 | tail -8 | tee gpurun_out/pytest_tc.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision f16x3 2>&1 | tail -1 > gpurun_out/bench_f16x3.json
python -c "
import json
d=json.load(open('gpurun_out/bench_f16x3.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['modes'].items()}); print({k[:40]:(v['n'],v['ms']) for k,v in d['roofline']['conv_by_layer'].items()}); print(d['e2e']); print({k:(v['n'],v['ms']) for k,v in d['roofline']['classes'].items()})"
HCF_BUILD_PROF=1 python -m hcflow_b200.build --force > /dev/null 2>&1; timeout 120 python tools/prof_chain.py f16x3 2>&1 | grep -E "hcf prof|precision" | tee gpurun_out/prof_chain5.log
python -m hcflow_b200.build --force > /dev/null 2>&1
