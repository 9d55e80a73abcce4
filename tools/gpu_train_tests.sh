#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "reference_training_loop or nll_gradients or training_step or inverse_path_l1 or reference_model_wrapper" 2>&1 | grep -v CUDAEvent | grep -E "^E|Error|error|assert|passed|failed" | head -20 | cut -c1-300 | tee gpurun_out/pytest_train.log
