#!/bin/bash
# ncu --set full of the dominant kernel at HEAD, at the bench's launch size (460 800 rows: one tile of 2 048 sequences)
cd "$(dirname "$0")/.."
mkdir -p /tmp/ncu
k=gtf_fwd_kernel
timeout 900 ncu --set full --clock-control none -k "regex:$k" -s 30 -c 24 -o /tmp/ncu/$k -f python tools/time_large.py --B 2048 --T 6 --steps 1 --precision 2 > /tmp/ncu/$k.out 2>&1
ncu -i /tmp/ncu/$k.ncu-rep --page raw --csv > /tmp/ncu/$k.csv 2>/dev/null
python - "$k" <<'PY'
import csv,sys,re
k=sys.argv[1]
rows=list(csv.reader(open('/tmp/ncu/%s.csv'%k)))
hdr=rows[0]; units=rows[1]
pat=re.compile(r'gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__pipe_tensor.*cycles_active.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread$|launch__grid_size|launch__block_size|smsp__issue_active.avg.pct|sm__throughput.avg.pct|lts__t_bytes.sum$|lts__throughput.avg.pct|l1tex__throughput.avg.pct|smsp__inst_executed.sum$|launch__shared_mem_per_block_dynamic|smsp__average_warps_issue_stalled.*_per_issue_active|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|Kernel Name|lts__t_sector_hit_rate.pct|sm__cycles_elapsed.avg$')
idx=[i for i,h in enumerate(hdr) if pat.search(h)]
ti=hdr.index('gpu__time_duration.sum')
scale={'ns':1e-3,'us':1.0,'usecond':1.0,'ms':1e3,'msecond':1e3,'nsecond':1e-3}.get(units[ti],1.0)
ni=hdr.index('Kernel Name')
big=[r for r in rows[2:] if float(r[ti].replace(',',''))*scale>150.0]
big=[r for r in big if '<1>' in r[ni]][:2]+[r for r in big if '<0>' in r[ni]][:1]
out=open('gpurun_out/r2c_ncu_big_%s.txt'%k,'w')
out.write('# ncu --set full --clock-control none -k regex:%s -s 30 -c 24 python tools/time_large.py --B 2048 --T 6 --steps 1 --precision 2 (launches over 460 800 rows only: the bench launch size)\n'%k)
for r in big:
    out.write('---- launch\n')
    for i in idx:
        out.write('%-90s %s %s\n'%(hdr[i],r[i],units[i]))
out.close()
print(k, 'big launches', len(big), 'of', len(rows)-2, [r[ni][:30] for r in big])
PY
grep -E "Kernel Name|gpu__time_duration|dram__bytes|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|issue_active" gpurun_out/r2c_ncu_big_$k.txt | head -14
