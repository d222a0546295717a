#!/bin/bash
# ncu --set full of the image-module kernels inside one encoder -> decoder forward + backward at 625 frames
# (gpurun -- 'bash tools/ncu_conv.sh'); the report is read here with `ncu -i ... --page raw --csv` (tools/ncu_conv_summary.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 140 ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'conv_gather|conv_scatter|conv_wgrad|dense_gemm|chan_reduce|bn_apply|bn_bwd_apply' --launch-count 40 \
  -o gpurun_out/r2_conv_full -f python tools/time_conv.py 625 > gpurun_out/r2_ncu_conv.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_conv.log; ls -la gpurun_out/r2_conv_full.ncu-rep
