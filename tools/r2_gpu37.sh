#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels: fused transition kernels, bfvi_mlp_*, a fused C3-dims step (tiled)
cd "$(dirname "$0")/.."
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py tests/test_gpu_mlp.py -x -q -k "not 19021 and not c4" > gpurun_out/r2_memcheck_fused.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_memcheck_fused.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_large.py -x -q -k "tiled and 2-4-2 or step_matches_oracle and c3_dims-2" > gpurun_out/r2_memcheck_step.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_memcheck_step.log
