mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:gemm_tf32_ts_kernel -s 30 -c 6 -o gpurun_out/s17_fwd -f python tools/time_large.py --B 256 --steps 1 > gpurun_out/s17_fwd.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:gemm_tf32_ts_kernel -s 2480 -c 8 -o gpurun_out/s17_bwd -f python tools/time_large.py --B 256 --steps 1 > gpurun_out/s17_bwd.log 2>&1
for f in fwd bwd; do ncu -i gpurun_out/s17_$f.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum > gpurun_out/s17_$f.csv 2>&1; done
tail -3 gpurun_out/s17_bwd.log
