#!/bin/bash
# time the product library under tuning knobs and every variant under tools/_variants
python tools/quick_time.py --tag base 2>&1 | tail -2
for c in 1 2 3 4; do BFVI_CHUNKS=$c python tools/quick_time.py --tag chunks$c 2>&1 | tail -2 | head -1; done
for f in tools/_variants/*.so; do
  [ -e "$f" ] || continue
  BFVI_LIB_PATH=$PWD/$f python tools/quick_time.py --tag $(basename $f .so) 2>&1 | tail -2
done
