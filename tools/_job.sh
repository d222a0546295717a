for i in 1 2; do timeout 120 python tools/quick_time.py --steps 10 --tag seg_auto 2>&1 | tail -2; done
BFVI_BWD_SEGMENTS=1 timeout 120 python tools/quick_time.py --steps 10 --tag seg_off 2>&1 | tail -2
timeout 120 python tools/quick_time.py --steps 10 --B 8192 --tag B8k 2>&1 | tail -2 | head -1
BFVI_BWD_SEGMENTS=1 timeout 120 python tools/quick_time.py --steps 10 --B 8192 --tag B8k_off 2>&1 | tail -2 | head -1
timeout 120 python tools/quick_time.py --steps 10 --B 2048 --tag B2k 2>&1 | tail -2 | head -1
BFVI_BWD_SEGMENTS=1 timeout 120 python tools/quick_time.py --steps 10 --B 2048 --tag B2k_off 2>&1 | tail -2 | head -1
