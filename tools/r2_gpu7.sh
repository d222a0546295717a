#!/bin/bash
# round 2 profiles: launch list of the bench command (short T) + ncu --set full of the fused kernels (text summaries only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/launches.csv python bench.py --workload c3 --batch 704 --seq-len 12 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_launches_bench.out 2>&1
python tools/launch_shares.py /tmp/ncu/launches.csv 40 > gpurun_out/r2_launches_bench_summary.txt
gzip -c /tmp/ncu/launches.csv > gpurun_out/r2_launches_bench.csv.gz
head -8 gpurun_out/r2_launches_bench_summary.txt
M='gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput|sm__pipe_tensor|sm__inst_executed_pipe_tensor|sm__warps_active|launch__registers_per_thread|launch__grid_size|smsp__issue_active|sm__throughput|l1tex__data_pipe_lsu_wavefronts_mem_shared|lts__t_bytes.sum|smsp__inst_executed.sum|launch__shared_mem|smsp__warp_issue_stalled.*_per_warp_active|sm__cycles_elapsed.avg '
for k in gtf_fwd_kernel gtf_bwd_kernel wgrad16_kernel; do
  timeout 600 ncu --set full --clock-control none -k regex:$k -s 2 -c 6 -o /tmp/ncu/$k -f python tools/time_large.py --B 704 --T 6 --steps 1 --precision 2 > /tmp/ncu/$k.out 2>&1
  ncu -i /tmp/ncu/$k.ncu-rep --page raw --csv > /tmp/ncu/$k.csv 2>/dev/null
  python - "$k" <<'PY'
import csv,sys,re
k=sys.argv[1]
rows=list(csv.reader(open('/tmp/ncu/%s.csv'%k)))
hdr=rows[0]; units=rows[1]
pat=re.compile(r'gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__pipe_tensor.*cycles_active.avg.pct|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct|launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__issue_active.avg.pct|sm__throughput.avg.pct|lts__t_bytes.sum$|smsp__inst_executed.sum$|launch__shared_mem_per_block_dynamic|smsp__average_warps_issue_stalled.*_per_issue_active|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|Kernel Name|Grid Size')
idx=[i for i,h in enumerate(hdr) if pat.search(h)]
out=open('gpurun_out/r2_ncu_%s.txt'%k,'w')
out.write('# ncu --set full --clock-control none -k regex:%s -s 2 -c 6 python tools/time_large.py --B 704 --T 6 --steps 1 --precision 2\n'%k)
for r in rows[2:]:
    out.write('---- launch\n')
    for i in idx:
        out.write('%-90s %s %s\n'%(hdr[i],r[i],units[i]))
out.close()
PY
  rm -f /tmp/ncu/$k.ncu-rep
  grep -E "gpu__time_duration|dram__bytes|tensor.*pct|Grid Size" gpurun_out/r2_ncu_$k.txt | head -12
done
du -sh gpurun_out
