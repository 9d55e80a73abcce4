#!/bin/bash
# gpurun_out/ of tools/gpu_r02_final.sh -> profiles/r02_* (tracked evidence)
set -e
cd "$(dirname "$0")/.."
g=gpurun_out; p=profiles
cp $g/bench_default.json $p/r02_bench_1gpu_f16x3.json
cp $g/bench_reference_arm.json $p/r02_bench_reference_arm.json
cp $g/bench_configs_f16x3.jsonl $p/r02_bench_other_configs_f16x3.jsonl
cp $g/bench_configs_f16.jsonl $p/r02_bench_other_configs_f16.jsonl
python - <<'PY'
import json
new = json.load(open('gpurun_out/parity_report.json'))
try:
    old = json.load(open('profiles/r02_parity_report.json'))
except Exception:
    old = {}
for k, v in old.items():          # the two NCCL tests only run on a 2-GPU box: keep their last result
    if 'nccl' in k and k not in new:
        new[k] = v
json.dump(new, open('profiles/r02_parity_report.json', 'w'), indent=1, sort_keys=True)
PY
cp $g/launches_f16x3.csv $p/r02_launches_f16x3.csv
cp $g/pytest_gpu.log $p/r02_pytest_gpu.log
cp $g/smoke.log $p/r02_smoke.log
cp $g/wait_profile_f16x3.log $p/r02_wait_profile_f16x3.log
cat $g/launch_times_rescaling.jsonl $g/launch_times_sr_x8.jsonl > $p/r02_launch_times_other_configs.jsonl
python tools/ncu_summary.py $g/prof_f16x3_chains.ncu-rep $p/r02_prof_chain_L0_f16x3_summary.csv > /dev/null
python tools/ncu_summary.py $g/prof_flowstep.ncu-rep $p/r02_prof_flowstep_f16x3_summary.csv > /dev/null
python tools/ncu_summary.py $g/prof_rescaling_dense.ncu-rep $p/r02_prof_rescaling_dense_chain_f16x3_summary.csv > /dev/null
python tools/ncu_summary.py $g/prof_rescaling_flowstep_fwd.ncu-rep $p/r02_prof_rescaling_flowstep_fwd_f16x3_summary.csv > /dev/null
ls $p | grep r02_ | wc -l
