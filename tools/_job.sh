for r in 1 2; do python tools/quick_time.py --steps 10 --tag rcp 2>&1 | tail -2; done
python tools/quick_time.py --B 100 --tag c1 2>&1 | tail -2
python -m pytest tests/test_gpu_step.py tests/test_gpu_model.py tests/test_gpu_random.py -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:chain_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_r1_pair_bwd25 python tools/quick_time.py --steps 1 --warmup 0 > gpurun_out/ncu_pair.log 2>&1
