"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys


def summarize(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1000, 'us': v, 'usecond': v, 'ms': v * 1000, 'msecond': v * 1000, 'nsecond': v / 1000}[row['Metric Unit']]
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
        tot[name] += v
        cnt[name] += 1
    return tot, cnt


if __name__ == '__main__':
    tot, cnt = summarize(sys.argv[1])
    T = sum(tot.values())
    print('%10s %5s %6s  kernel' % ('total_us', 'count', 'share'))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('%10.1f %5d %5.1f%%  %s' % (v, cnt[k], 100 * v / T, k[:90]))
    print('%10.1f total' % T)
