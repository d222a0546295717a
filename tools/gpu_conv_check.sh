#!/bin/bash
# B200 check of the image-module kernels (gpurun -- 'bash tools/gpu_conv_check.sh'): their parity tests + the Weizmann
# golden through them, kernel timings next to torch / cuDNN, the C4 bench line, then the whole GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_conv.py tests/test_gpu_weizmann.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_conv_tests.log 2>&1; echo "conv tests rc=$?"; tail -15 gpurun_out/r2_conv_tests.log
if ! tail -1 gpurun_out/r2_conv_tests.log | grep -q " passed" || tail -1 gpurun_out/r2_conv_tests.log | grep -q failed; then
  BFVI_IMAGE_KERNELS=conv timeout 100 python -m pytest tests/test_gpu_conv.py tests/test_gpu_weizmann.py -q --tb=line -p no:cacheprovider -k "modules or weizmann" > gpurun_out/r2_conv_tests_convonly.log 2>&1; echo "conv-only rc=$?"; tail -8 gpurun_out/r2_conv_tests_convonly.log
fi
timeout 60 python tools/time_conv.py > gpurun_out/r2_time_conv.json 2> gpurun_out/r2_time_conv.err; echo "time_conv rc=$?"; cut -c1-600 gpurun_out/r2_time_conv.json
timeout 100 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/r2_c4_conv.json 2> gpurun_out/r2_c4_conv.err; echo "c4 rc=$?"; cut -c1-300 gpurun_out/r2_c4_conv.json
timeout 200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_conv_full_tests.log 2>&1; echo "full suite rc=$?"; tail -3 gpurun_out/r2_conv_full_tests.log
