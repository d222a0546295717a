#!/bin/bash
cd "$(dirname "$0")/.."
timeout 120 python -m pytest tests/test_gpu_fused.py -q -x --timeout 60 2>&1 | tail -6
[ ${PIPESTATUS[0]} -eq 0 ] || exit 1
timeout 200 python -m pytest tests/test_gpu_large.py -q -x --timeout 100 -k "step_matches_oracle or tiled" 2>&1 | tail -4
BFVI_FUSED_DBG=1 timeout 60 python -m pytest tests/test_gpu_fused.py -q -x -k "lattice_forward and 512 and 19021" -s 2>&1 | grep -E "dbg|issuer" | head -4
timeout 120 python tools/time_large.py --B 256 --T 100 --steps 3 --precision 2 2>&1 | tail -1
timeout 200 python tools/time_large.py --B 1024 --T 50 --steps 2 --precision 2 --batch-tile 512 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g4_launches.csv python tools/time_large.py --B 256 --T 20 --steps 1 --precision 2 > gpurun_out/r2_g4_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/r2_g4_launches.csv 14
