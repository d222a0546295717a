#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
BFVI_FUSED_ABL=64 timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
timeout 600 python tools/probe_fused_ablate.py 0 64 65 66 > gpurun_out/r2_fused_ablate2.txt 2>&1
cat gpurun_out/r2_fused_ablate2.txt
for abl in 0 64; do BFVI_FUSED_ABL=$abl timeout 300 python tools/time_large.py --B 704 --T 40 --steps 2 2>&1 | grep ms/step; done
BFVI_FUSED_ABL=64 timeout 600 python -m pytest tests/test_gpu_large.py -x -q -k "step_matches_oracle" 2>&1 | tail -2
