#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "fused_flowstep_kernel" 2>&1 | tail -5 | tee gpurun_out/pytest_flowstep.log
if grep -q "failed\|error" gpurun_out/pytest_flowstep.log; then exit 1; fi
timeout -k 5 700 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-modes --no-cpu-baseline 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'])
for k,v in d['roofline']['conv_by_layer'].items(): print(k, v)
PY
HCFLOW_LIB=$PWD/hcflow_b200/prof/libhcflow_b200_prof.so HCF_TC_PROF=1 timeout -k 5 300 python tools/prof_chain.py f16x3 2> gpurun_out/wait_profile_f16x3.log
grep flowstep gpurun_out/wait_profile_f16x3.log
timeout -k 5 1700 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
