"""Per-launch summary of an `ncu --set full` report of the image-module kernels (tools/ncu_conv.sh):
`python tools/ncu_conv_summary.py gpurun_out/r2_conv_full.ncu-rep > profiles/r2_conv_ncu_summary.txt`"""
import csv
import io
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
           'launch__registers_per_thread', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv', '--metrics', ','.join(METRICS)],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
head, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(head)}
print('%-34s %-14s %8s %8s %8s %6s %6s %6s %6s %6s %6s %6s %5s %6s %6s %6s' % (
    'kernel', 'grid', 'us', 'rd MB', 'wr MB', 'sm%', 'fma%', 'lsu%', 'issue%', 'warps%', 'L1hit', 'L2hit', 'regs', 'st_lsb', 'st_ssb', 'st_mio'))
for r in body:
    name = r[col['Kernel Name']].replace('void ', '').split('(')[0]
    g = lambda m: float(r[col[m]]) if m in col and r[col[m]] not in ('', 'n/a') else float('nan')
    print('%-34s %-14s %8.1f %8.1f %8.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f %5d %6.2f %6.2f %6.2f' % (
        name, r[col['Grid Size']].replace(' ', ''), g(METRICS[0]), g(METRICS[1]), g(METRICS[2]), g(METRICS[3]), g(METRICS[4]),
        g(METRICS[5]), g(METRICS[6]), g(METRICS[7]), g(METRICS[8]), g(METRICS[9]), int(g(METRICS[10])), g(METRICS[11]),
        g(METRICS[12]), g(METRICS[13])))
