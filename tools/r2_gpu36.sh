#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_large.py tests/test_gpu_model.py -x -q 2>&1 | tail -2
timeout 300 python tools/time_large.py --B 2048 --T 40 --steps 3 2>&1 | grep -E "ms/step"
