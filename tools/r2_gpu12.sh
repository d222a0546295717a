#!/bin/bash
# where does a C3-dims step go?  passes switched off one at a time, side stream on / off, eager / graph
cd "$(dirname "$0")/.."
O=gpurun_out/r2_pass_breakdown.txt; : > $O
run() { echo "== $*" >> $O; env "${@:2}" timeout 300 python tools/time_large.py --B 704 --T 40 --steps 2 $1 2>&1 | grep -E "ms/step|Error|error" >> $O; }
run "" X=1
run "" BFVI_LARGE_SIDE=0
run "--f-mult 0" X=1
run "--s-mult 0" X=1
run "--match-mult 0" X=1
run "--graph" X=1
run "--precision 0" X=1
run "--fwd-only" X=1
cat $O
