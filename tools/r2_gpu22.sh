#!/bin/bash
cd "$(dirname "$0")/.."
BFVI_FUSED_DBG=1 timeout 120 python -m pytest tests/test_gpu_fused.py -q -x -k "lattice_forward and 512 and 19021" -s 2>&1 | grep -E "dbg|issuer|head row" | head -4
BFVI_FUSED_DBG=1 BFVI_FUSED_ABL=64 timeout 120 python -m pytest tests/test_gpu_fused.py -q -x -k "lattice_forward and 512 and 19021" -s 2>&1 | grep -E "dbg|issuer|head row" | head -4
