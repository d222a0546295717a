#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g6_c3.json 2> gpurun_out/r2_g6_c3.err
tail -3 gpurun_out/r2_g6_c3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_g6_c3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','dtype','gpu_launches','clocks')})
print(d['e2e']); print(d['roofline']); print(json.dumps(d['kernel_probe'],indent=1)); print(d['dispatch'])
PY
