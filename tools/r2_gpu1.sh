#!/bin/bash
# round 2, GPU call 1: parity of the bench-path kernel variants + first C3 bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_g1_smi.txt
timeout 1500 python -m pytest tests/test_gpu_variants.py -x -q > gpurun_out/r2_g1_variants.log 2>&1
echo "variants rc=$?" >> gpurun_out/r2_g1_variants.log
timeout 600 python bench.py --workload c3 --batch 512 --steps 3 --warmup 3 --precision tf32x3 --no-cpu-baseline > gpurun_out/r2_g1_c3_x3.json 2> gpurun_out/r2_g1_c3_x3.err
timeout 600 python bench.py --workload c3 --batch 512 --steps 3 --warmup 3 --precision fused > gpurun_out/r2_g1_c3_tf32.json 2> gpurun_out/r2_g1_c3_tf32.err
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 > gpurun_out/r2_g1_c2.json 2> gpurun_out/r2_g1_c2.err
timeout 900 python bench.py --impl reference --workload c3 --steps 3 --warmup 1 > gpurun_out/r2_g1_c3_ref.json 2> gpurun_out/r2_g1_c3_ref.err
tail -3 gpurun_out/r2_g1_variants.log
cat gpurun_out/r2_g1_c3_x3.json gpurun_out/r2_g1_c3_tf32.json | cut -c1-600
