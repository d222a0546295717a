#!/bin/bash
# end-of-round measurements at HEAD: bench lines (C3 default, reference arm, C2) + launch list of the bench command
cd "$(dirname "$0")/.."
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_c3.json 2> gpurun_out/r2_final_c3.err
tail -2 gpurun_out/r2_final_c3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_c3_ref.json 2> gpurun_out/r2_final_c3_ref.err
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 > gpurun_out/r2_final_c2.json 2> gpurun_out/r2_final_c2.err
python - <<'PY'
import json
for f in ('r2_final_c3', 'r2_final_c3_ref', 'r2_final_c2'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')})
        print('  e2e', d.get('e2e')); print('  roofline', d.get('roofline')); print('  cpu', d.get('cpu_baseline')); print('  probe', d.get('kernel_probe'))
    except Exception as e:
        print(f, 'ERR', e)
PY
mkdir -p /tmp/ncu
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/launches.csv python bench.py --workload c3 --batch 1024 --seq-len 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_launches.out 2>&1
python tools/launch_shares.py /tmp/ncu/launches.csv 40 > gpurun_out/r2_final_launches_summary.txt
gzip -c /tmp/ncu/launches.csv > gpurun_out/r2_final_launches.csv.gz
head -24 gpurun_out/r2_final_launches_summary.txt
