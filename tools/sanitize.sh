#!/bin/bash
# compute-sanitizer over one small-dim and one large-dim training step + forward (development aid)
for tool in memcheck racecheck; do
  echo "== $tool: small-dim family"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_step.py -q -x -k "spirals_ragged or forward_only" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Race|Invalid|hazard" | head -8
  echo "== $tool: large-dim family"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_large.py -q -x -k "odd or default32" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Race|Invalid|hazard" | head -8
done
echo "== memcheck: losses + batch preparation"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_losses.py tests/test_gpu_multiseq.py -q -x 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head -8
