#!/bin/bash
# compute-sanitizer memcheck over whole passes at golden size (every kernel of the inference and training paths)
mkdir -p gpurun_out
timeout -k 5 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 1200 -k "(reverse_matches_reference_golden) or (sr_forward_matches_reference_golden and sr_x4) or rescaling_forward_matches_reference_golden or (tensor_core_modes_match and f16x3) or uint8 or (nll_gradients and sr_x4 and regular) or tiled_inference or psnr" 2>&1 | grep -v CUDAEvent | tail -12 | cut -c1-300 | tee gpurun_out/sanitizer_memcheck_passes.log
