#!/bin/bash
# build the CUDA library locally (so the fresh .so travels), then run a command on the GPU box
# usage: tools/grun.sh <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" >/dev/null
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
