for r in 1 2; do
python tools/quick_time.py --steps 10 --tag pf 2>&1 | tail -2
BFVI_LIB_PATH=$PWD/tools/_variants/libbfvi_nopf.so python tools/quick_time.py --steps 10 --tag nopf 2>&1 | tail -2
BFVI_ZSPLIT_BWD=0 python tools/quick_time.py --steps 10 --tag pf_zb0 2>&1 | tail -2
done
python tools/quick_time.py --B 100 --tag c1 2>&1 | tail -2
BFVI_LIB_PATH=$PWD/tools/_variants/libbfvi_nopf.so python tools/quick_time.py --B 100 --tag c1nopf 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:zsplit -c 4 -o gpurun_out/zsplit_pf python tools/quick_time.py --steps 1 --warmup 0 > gpurun_out/ncu_zsplit.log 2>&1
