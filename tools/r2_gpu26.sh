#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -k "long_chain or step_matches_oracle" -s 2>&1 | grep -E "errors|condition|passed|failed|Error|assert" | cut -c1-600 | head -24
