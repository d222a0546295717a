#!/bin/bash
cd "$(dirname "$0")/.."
echo "== skip wgrad (K2a only)"; BFVI_DBG_SKIP_WGRAD=1 timeout 40 python -m pytest tests/test_gpu_fused.py -q -x -k "lattice_backward and 300" 2>&1 | tail -3; echo "rc=$?"
echo "== skip bwd (wgrad only)"; BFVI_DBG_SKIP_BWD=1 timeout 40 python -m pytest tests/test_gpu_fused.py -q -x -k "lattice_backward and 300" 2>&1 | tail -3; echo "rc=$?"
nvidia-smi --query-gpu=utilization.gpu --format=csv,noheader
