#!/bin/bash
cd "$(dirname "$0")/.."
BFVI_FUSED_DBG=1 timeout 60 python -m pytest tests/test_gpu_fused.py -q -x -k "lattice_forward and 512" -s 2>&1 | grep -E "dbg|issuer" | head -12
