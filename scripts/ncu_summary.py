"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel name."""
import collections, csv, re, sys

def main(path, skip_first=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    tot = 0.0
    for i, row in enumerate(csv.DictReader(lines)):
        if i < skip_first:
            continue
        name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', row['Kernel Name'])
        name = re.sub(r'\(.*$', '', name)
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v; tot += v
    print('%12s %6s %7s  kernel' % ('total_us', 'count', 'share'))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%12.1f %6d %6.1f%%  %s' % (t, c, 100 * t / tot, k))
    print('%12.1f total' % tot)

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
