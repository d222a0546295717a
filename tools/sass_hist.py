"""Opcode histogram of an `ncu --page source --csv` export (executed warp-instructions
and stall samples per SASS mnemonic).  usage: python tools/sass_hist.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
i_src, i_exec, i_samp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ex, sm = collections.Counter(), collections.Counter()
section, want = 0, int(sys.argv[3]) if len(sys.argv) > 3 else 0
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        section += 1
        continue
    if len(r) <= i_exec or not r[i_exec].isdigit() or section != want:
        continue
    toks = r[i_src].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('MUFU', 'LDS', 'STS', 'LDG', 'STG', 'IMAD', 'SHFL')) and '.' in op else '')
    ex[op] += int(r[i_exec] or 0)
    sm[op] += int(r[i_samp] or 0)
tot, tots = sum(ex.values()), sum(sm.values())
print('total executed warp-instr %d, static SASS lines %d, samples %d' % (tot, len(rows) - 2, tots))
for op, n in ex.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print('%-14s %12d %5.1f%%   samples %5.1f%%' % (op, n, 100.0 * n / tot, 100.0 * sm[op] / max(tots, 1)))
