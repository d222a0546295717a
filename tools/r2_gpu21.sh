#!/bin/bash
# 2 GPUs: NCCL parity test + C3 bench lines (weak and strong)
cd "$(dirname "$0")/.."
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parallel.py -x -q 2>&1 | tail -4
for sc in weak strong; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --scaling $sc > gpurun_out/r2_c3_n2_$sc.json 2> gpurun_out/r2_c3_n2_$sc.err
tail -2 gpurun_out/r2_c3_n2_$sc.err; python - $sc <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_c3_n2_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','clocks')}); print(d['e2e']); print(d['config']['workload'][:80], d['config']['global_batch'])
PY
done
