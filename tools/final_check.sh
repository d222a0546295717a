mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s13_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/s13_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s13_smoke.log 2>&1; tail -1 gpurun_out/s13_smoke.log
timeout 300 python bench.py > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/s13_bench.json
timeout 200 python tools/bench_configs.py c3 --steps 5 > gpurun_out/s13_c3.json 2> gpurun_out/s13_c3.err; echo "c3 rc=$?"; head -c 300 gpurun_out/s13_c3.json
timeout 120 python tools/time_large.py --B 256 --steps 5 --graph 2>&1 | tail -1
