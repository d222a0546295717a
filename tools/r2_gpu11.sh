#!/bin/bash
cd "$(dirname "$0")/.."
ls oracle/_ref oracle/_ref/models | head -20
timeout 600 python -m pytest tests/test_reference_dropin.py -x -q 2>&1 | tail -15
(time timeout 900 python bench.py --impl reference --steps 2 --warmup 1) 2>&1 | grep -o '"cpu_baseline.*' | cut -c1-600
