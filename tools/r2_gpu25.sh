#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
