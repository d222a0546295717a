#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for prec in 0 2; do
  timeout 300 python tools/time_large.py --B 256 --T 100 --steps 3 --precision $prec 2>&1 | tail -2
done > gpurun_out/r2_g3_time.log 2>&1
timeout 300 python tools/time_large.py --B 1024 --T 50 --steps 2 --precision 2 2>&1 | tail -1 >> gpurun_out/r2_g3_time.log
cat gpurun_out/r2_g3_time.log
# launch list of one fused step (T=20): shares per kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_g3_launches.csv python tools/time_large.py --B 256 --T 20 --steps 1 --precision 2 > gpurun_out/r2_g3_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/r2_g3_launches.csv 25 | tee gpurun_out/r2_g3_shares.txt
