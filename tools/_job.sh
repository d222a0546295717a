python -m pytest tests/test_gpu_losses.py tests/test_gpu_multiseq.py tests/test_gpu_weizmann.py -x -q 2>&1 | tail -5
python tools/bench_streaming.py 2>&1 | tee gpurun_out/streaming.jsonl | cut -c1-400
