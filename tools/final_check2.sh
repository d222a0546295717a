mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s20_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/s20_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s20_smoke.log 2>&1; tail -1 gpurun_out/s20_smoke.log
timeout 200 python tools/bench_configs.py c3 --steps 5 > gpurun_out/s20_c3.json 2> gpurun_out/s20_c3.err; echo "c3 rc=$?"; cut -c1-300 gpurun_out/s20_c3.json
timeout 100 python tools/bench_configs.py c3 --batch 2048 --T 25 --steps 5 > gpurun_out/s20_c3_b2048.json 2>> gpurun_out/s20_c3.err; cut -c1-200 gpurun_out/s20_c3_b2048.json
