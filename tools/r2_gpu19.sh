#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/probe_fused_ablate.py 0 2>&1 | tail -1
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_g19_c3.json 2> gpurun_out/r2_g19_c3.err
tail -3 gpurun_out/r2_g19_c3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_g19_c3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print(d['e2e']); print(d['roofline']); print(d['cpu_baseline']); print(d['dispatch'])
PY
