#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -q --timeout 120 > gpurun_out/r2_g2_fused.log 2>&1
echo "fused rc=$?" >> gpurun_out/r2_g2_fused.log
tail -30 gpurun_out/r2_g2_fused.log
timeout 900 python -m pytest tests/test_gpu_large.py -q --timeout 300 -k "step_matches_oracle" -s > gpurun_out/r2_g2_large.log 2>&1
echo "large rc=$?" >> gpurun_out/r2_g2_large.log
grep -E "large-step|passed|failed|Error|rc=" gpurun_out/r2_g2_large.log | tail -20
