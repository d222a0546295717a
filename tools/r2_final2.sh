#!/bin/bash
# end-of-round measurements at HEAD (after the one-tile step): C3 default with the cpu baseline, C5 inference line
cd "$(dirname "$0")/.."
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final2_c3.json 2> gpurun_out/r2_final2_c3.err
tail -2 gpurun_out/r2_final2_c3.err
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r2_final2_c5.json 2> gpurun_out/r2_final2_c5.err
tail -2 gpurun_out/r2_final2_c5.err
python - <<'PY'
import json
for f in ('r2_final2_c3', 'r2_final2_c5'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('metric','value','ms_per_step','gpu_launches','clocks')})
        print('  e2e', d.get('e2e')); r=d.get('roofline'); print('  roofline', r and {k:r.get(k) for k in ('frac','achieved','ms_per_launch','rows_per_launch','traffic')}, r and r.get('whole_step')); print('  cpu', d.get('cpu_baseline')); print('  dispatch', d.get('dispatch'))
    except Exception as e:
        print(f, 'ERR', e)
PY
