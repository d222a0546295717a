#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/probe_fused_ablate.py > gpurun_out/r2_fused_ablate.txt 2>&1
cat gpurun_out/r2_fused_ablate.txt
