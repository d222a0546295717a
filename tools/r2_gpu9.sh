#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/probe_step_errors.py 5 1 25 > gpurun_out/r2_probe_step_errors.txt 2>&1
cat gpurun_out/r2_probe_step_errors.txt
