#!/bin/bash
# one GPU call: tests + smoke + quick timing.  usage: tools/gpu_check.sh [extra pytest args]
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -25)
(timeout 120 python __graft_entry__.py 2>&1 | tail -3)
timeout 200 python tools/quick_time.py --steps 5 --warmup 2 2>&1 | tail -2
