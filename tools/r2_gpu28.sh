#!/bin/bash
cd "$(dirname "$0")/.."
BFVI_MLP_CHUNK=32 timeout 900 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_model.py tests/test_gpu_weizmann.py -x -q 2>&1 | tail -2
timeout 300 python tools/time_large.py --B 2048 --T 40 --steps 2 2>&1 | grep -E "ms/step"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g28_c3.json 2> gpurun_out/r2_g28_c3.err
tail -2 gpurun_out/r2_g28_c3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_g28_c3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['dispatch']); print(d['roofline']['frac'], d['roofline']['whole_step']['frac'])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
