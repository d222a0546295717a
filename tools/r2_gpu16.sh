#!/bin/bash
cd "$(dirname "$0")/.."
for s in 2 3 4 6 8 9; do echo "wgrad stages $s"; BFVI_WGRAD_STAGES=$s timeout 300 python tools/probe_fused_ablate.py 0 2>&1 | tail -1; done
