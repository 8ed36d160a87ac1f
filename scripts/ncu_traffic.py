"""profiles/r02_kernel_traffic.json from `ncu --set full` exports (--page raw --csv) of the CURRENT build:
per kernel (template name without arguments) the mean dram read + write bytes per launch, plus the sum over
the E-step kernels as 'arhmm_estep'.  bench.py reads the file for `roofline.traffic`.

    python scripts/ncu_traffic.py cae.raw.csv [hmm.raw.csv] > profiles/r02_kernel_traffic.json
"""
import collections
import csv
import json
import re
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(paths):
    tot, cnt = collections.defaultdict(float), collections.Counter()
    estep = 0.0
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name')
        per_kernel = collections.defaultdict(list)
        for r in rows[2:]:
            name = re.sub(r'\(.*', '', r[ik]).replace('void ', '').replace('<unnamed>::', '').replace(' ', '')
            b = float(r[ir].replace(',', '')) * UNIT[units[ir]] + float(r[iw].replace(',', '')) * UNIT[units[iw]]
            per_kernel[name].append(b)
        for k, v in per_kernel.items():
            tot[k] += sum(v)
            cnt[k] += len(v)
        if any(k.startswith(('scan2_kernel', 'emission')) for k in per_kernel):
            # one E-step = one launch of each distinct E-step kernel (x the trial groups): sum of the means x count ratio
            n_scan = max(len(v) for k, v in per_kernel.items() if k.startswith('scan2_kernel'))
            n_groups = [len(v) for k, v in per_kernel.items() if k.startswith(('scan2_kernel', 'emission'))]
            estep = sum(sum(v) for k, v in per_kernel.items() if k.startswith(('scan2_kernel', 'emission')))
            estep /= max(1, min(n_groups)) if False else 1.0
            estep = estep / float(sys.argv[-1]) if sys.argv[-1].isdigit() else estep
    out = {k: tot[k] / cnt[k] for k in sorted(tot)}
    if estep:
        out['arhmm_estep'] = estep
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main([p for p in sys.argv[1:] if not p.isdigit()])
