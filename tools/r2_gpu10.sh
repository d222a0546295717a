#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_reference_dropin.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_large.py -q -k "step_matches_oracle" -s 2>&1 | grep -E "errors|condition|passed|failed|Error|assert" | head -20
timeout 600 python -m pytest tests/test_gpu_random.py tests/test_gpu_weizmann.py -q 2>&1 | tail -15
(time timeout 900 python bench.py --impl reference --steps 2 --warmup 1) 2>&1 | cut -c1-1500 | tail -8
