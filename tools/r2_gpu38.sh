#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/r2_final_c4.json 2> gpurun_out/r2_final_c4.err
tail -3 gpurun_out/r2_final_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_c4.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('metric','value','ms_per_step','clocks')}); print(d['e2e']); print(d['cpu_baseline']); print(d['dispatch'])
PY
