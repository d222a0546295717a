for i in 1 2; do timeout 120 python tools/quick_time.py --steps 10 --tag seg_auto 2>&1 | tail -2; done
BFVI_FWD_SEGMENTS=1 timeout 120 python tools/quick_time.py --steps 10 --tag fwdseg_off 2>&1 | tail -2
BFVI_FWD_SEGMENTS=4 timeout 120 python tools/quick_time.py --steps 10 --tag fwdseg4 2>&1 | tail -2
BFVI_FWD_SEGMENTS=16 timeout 120 python tools/quick_time.py --steps 10 --tag fwdseg16 2>&1 | tail -2
BFVI_COOPERATIVE=0 timeout 120 python tools/quick_time.py --steps 10 --tag multilaunch 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_model.py tests/test_gpu_random.py -x -q 2>&1 | tail -2
