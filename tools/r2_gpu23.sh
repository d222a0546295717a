#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_large.py -x -q -k "forward" 2>&1 | tail -3
timeout 300 python tools/time_forward.py 2>&1 | tail -3
timeout 300 python tools/time_forward.py --B 256 --T 1000 --K 200 --steps 1 2>&1 | tail -3
timeout 300 python tools/probe_fused_ablate.py 0 2>&1 | tail -1
