#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -2
timeout 300 python tools/probe_fused_ablate.py 0 191 63 > gpurun_out/r2_fused_ablate4.txt 2>&1; cat gpurun_out/r2_fused_ablate4.txt
timeout 300 python tools/time_large.py --B 704 --T 40 --steps 2 2>&1 | grep ms/step
timeout 300 python tools/time_large.py --B 1024 --T 40 --steps 2 2>&1 | grep ms/step
BFVI_FUSED_DBG=1 timeout 120 python -m pytest tests/test_gpu_fused.py -q -x -k "lattice_forward and 512 and 19021" -s 2>&1 | grep -E "dbg|issuer" | head -4
