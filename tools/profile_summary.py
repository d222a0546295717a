"""Text summary of an .ncu-rep (key metrics per kernel + SASS opcode histogram) for profiles/.
usage: python tools/profile_summary.py report.ncu-rep > profiles/name.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__pcsamp_warps_issue_stalled_wait_not_issued',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard_not_issued',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard_not_issued',
        'smsp__pcsamp_warps_issue_stalled_branch_resolving_not_issued',
        'smsp__pcsamp_warps_issue_stalled_no_instructions_not_issued',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle_not_issued',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall_not_issued',
        'smsp__pcsamp_warps_issue_stalled_barrier_not_issued',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle_not_issued',
        'smsp__pcsamp_warps_issue_stalled_selected']
print('source report: %s  (ncu --set full --clock-control none --import-source on)' % rep.split('/')[-1])
for r in rows[2:]:
    print('=' * 100)
    print(r[hdr.index('Kernel Name')])
    for w in WANT:
        if w in hdr:
            print('  %-70s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
open('/tmp/_src.csv', 'w').write(src)
n_sections = src.count('"Kernel Name"')
n = len(rows) - 2                       # kernels in the report; the source export may repeat each one
per = max(n_sections // max(n, 1), 1)
for k in range(n):
    print('=' * 100)
    print('SASS opcode histogram, kernel #%d' % k)
    print(subprocess.run([sys.executable, __file__.replace('profile_summary', 'sass_hist'), '/tmp/_src.csv', '24', str(k * per)],
                         capture_output=True, text=True).stdout)
