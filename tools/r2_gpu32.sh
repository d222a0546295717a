#!/bin/bash
cd "$(dirname "$0")/.."
BFVI_LIB_PATH=$PWD/tools/_variants/libbfvi_ablate.so timeout 300 python tools/probe_fused_ablate.py 0 256 1 2>&1 | tail -4
