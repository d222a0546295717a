#!/bin/bash
# End-of-round verification on a B200 box (gpurun -- 'bash tools/final_check.sh'): the GPU test suite, smoke(), the bench
# line of both arms at the default (C3) workload.  Build first, here, with `python -c "import __graft_entry__ as g; g.build()"`.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/final_bench_ref.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final_bench.json
