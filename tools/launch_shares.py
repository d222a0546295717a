"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_shares.py <launches.csv> [top]"""
import collections
import csv
import re
import sys


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%-56s n=%5d  %10.1f us  %5.1f%%  avg %6.1f us' % (k[:56], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    print('total %.1f us over %d launches' % (tot, sum(v[0] for v in agg.values())))


if __name__ == '__main__':
    main()
