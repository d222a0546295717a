#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/time_large.py --B 2048 --T 1000 --steps 2 2>&1 | grep -E "ms/step" | cut -c1-160
for l in 1 2; do BFVI_TILE_LANES=$l timeout 300 python tools/time_large.py --B 2048 --T 1000 --steps 2 --batch-tile 1024 2>&1 | grep -E "ms/step" | cut -c1-160; done
