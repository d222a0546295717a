#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final3_c3.json 2> gpurun_out/r2_final3_c3.err
tail -2 gpurun_out/r2_final3_c3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final3_c3.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('metric','value','ms_per_step','gpu_launches','clocks')})
print('  e2e', d.get('e2e')); r=d.get('roofline'); print('  roofline', {k:r.get(k) for k in ('frac','achieved','ms_per_launch','rows_per_launch','traffic')}, r.get('whole_step')['frac']); print('  cpu', d.get('cpu_baseline')); print(' probe', {k:(round(v['ms'],4), round(v['frac'],4)) for k,v in d['kernel_probe'].items()})
PY
