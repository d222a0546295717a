#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
timeout 300 python tools/probe_f16_range.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_large.py -x -q -k "step_matches_oracle or long_chain or tiled" 2>&1 | tail -2
timeout 300 python tools/probe_fused_ablate.py 0 2>&1 | tail -1
timeout 300 python tools/time_large.py --B 2048 --T 40 --steps 2 2>&1 | grep -E "ms/step"
