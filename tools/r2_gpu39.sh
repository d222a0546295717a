#!/bin/bash
# the named per-GPU shard of C3: B = 8 192 sequences x T = 1 000 in ONE step() call (four batch tiles of 2 048)
cd "$(dirname "$0")/.."
timeout 700 python bench.py --batch 8192 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_c3_b8192.json 2> gpurun_out/r2_c3_b8192.err
tail -3 gpurun_out/r2_c3_b8192.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c3_b8192.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['dispatch']); print(d['config']['workload'][:90])
PY
