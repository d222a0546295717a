#!/bin/bash
cd "$(dirname "$0")/.."
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) 2>&1 | tail -7
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
