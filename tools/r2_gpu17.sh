#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/probe_fused_ablate.py 0 128 191 > gpurun_out/r2_fused_ablate3.txt 2>&1; cat gpurun_out/r2_fused_ablate3.txt
