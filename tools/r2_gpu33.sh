#!/bin/bash
# N GPUs: C3 bench line (weak)
cd "$(dirname "$0")/.."
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2_c3_n${N}_weak.json 2> gpurun_out/r2_c3_n${N}_weak.err
tail -2 gpurun_out/r2_c3_n${N}_weak.err; python - $N <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_c3_n%s_weak.json'%sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','clocks')}); print(d['e2e']); print(d['config']['global_batch'])
PY
