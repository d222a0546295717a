#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_large.py -x -q 2>&1 | tail -2
timeout 300 python tools/probe_fused_ablate.py 0 2>&1 | tail -1
BFVI_LIB_PATH=$PWD/tools/_variants/libbfvi_ablate.so timeout 300 python tools/probe_fused_ablate.py 0 1 2 4 8 16 32 63 191 > gpurun_out/r2_fused_ablate5.txt 2>&1; cat gpurun_out/r2_fused_ablate5.txt
for b in 704 1024; do timeout 300 python tools/time_large.py --B $b --T 40 --steps 2 2>&1 | grep ms/step; done
