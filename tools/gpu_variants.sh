#!/bin/bash
# time the product library and every tuning variant under tools/_variants (one GPU call)
python tools/quick_time.py --tag base 2>&1 | tail -2
for f in tools/_variants/*.so; do
  BFVI_LIB_PATH=$PWD/$f python tools/quick_time.py --tag $(basename $f .so) 2>&1 | tail -2
done
