#!/bin/bash
# round 2: FP16 dynamic-range probe of the fused kernels + cycle counters of the fused forward
cd "$(dirname "$0")/.."
timeout 300 python tools/probe_f16_range.py > gpurun_out/r2_probe_f16_range.txt 2>&1
cat gpurun_out/r2_probe_f16_range.txt
BFVI_FUSED_DBG=1 timeout 120 python -m pytest tests/test_gpu_fused.py -q -x -k "lattice_forward and 512 and 19021" -s 2>&1 | grep -E "dbg|issuer" | head -4
timeout 300 python -m pytest tests/test_gpu_large.py -q -x -k "step_matches_oracle and c3_dims" -s 2>&1 | grep -E "errors|passed|failed"
