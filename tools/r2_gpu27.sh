#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_large.py -x -q -k "tiled" 2>&1 | tail -3
timeout 300 python tools/time_large.py --B 1024 --T 40 --steps 2 2>&1 | grep ms/step
for l in 1 2; do BFVI_TILE_LANES=$l timeout 300 python tools/time_large.py --B 1024 --T 40 --steps 2 --batch-tile 512 2>&1 | grep -E "ms/step|dispatch" | cut -c1-200; done
for l in 1 2; do BFVI_TILE_LANES=$l timeout 300 python tools/time_large.py --B 2048 --T 40 --steps 2 --batch-tile 1024 2>&1 | grep -E "ms/step" | cut -c1-200; done
BFVI_TILE_LANES=2 timeout 300 python tools/time_large.py --B 2048 --T 40 --steps 2 --batch-tile 512 2>&1 | grep -E "ms/step" | cut -c1-200
