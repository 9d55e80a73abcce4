#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_f16x3.json
python -c "
import json
d=json.load(open('gpurun_out/bench_f16x3.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['modes'].items()}); print({k[:40]:(v['n'],v['ms']) for k,v in d['roofline']['conv_by_layer'].items()}); print(d['e2e']); print({k:(v['n'],v['ms']) for k,v in d['roofline']['classes'].items()}); print(d['config']['conv_kernels'], d['launches_per_step'])"
