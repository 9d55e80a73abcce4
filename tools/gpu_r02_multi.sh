#!/bin/bash
# multi-GPU: NCCL batch-NLL test (2 ranks) + bench with the configs[4] block
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -k "nccl" 2>&1 | tail -5 | tee gpurun_out/pytest_nccl.log
fi
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_${N}gpu.err | tail -1 > gpurun_out/bench_${N}gpu.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print({k:d[k] for k in ['value','ms_per_step','n_gpus','e2e']})
print(json.dumps(d.get('configs4'), indent=1))
PY
tail -3 gpurun_out/bench_${N}gpu.err
