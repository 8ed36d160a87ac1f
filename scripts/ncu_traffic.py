"""profiles/r02_kernel_traffic.json from `ncu --set full` exports (--page raw --csv) of the CURRENT build:
per kernel (template name without arguments, spaces removed) the mean dram read + write bytes per launch and,
for an export of ONE ARHMM E-step, the sum over its kernels as 'arhmm_estep' plus the executed warp instructions and
tensor-pipe cycles of its two kernels ('arhmm_estep_bounds': the issue-slot and tensor floors bench.py reports next
to the HBM roofline).  bench.py reads the file for `roofline.traffic`.

    python scripts/ncu_traffic.py cae.raw.csv [--estep hmm.raw.csv] > profiles/r02_kernel_traffic.json
"""
import collections
import csv
import json
import re
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name')
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[ik]).replace('void ', '').replace('<unnamed>::', '').replace(' ', '')
        yield name, float(r[ir].replace(',', '')) * UNIT[units[ir]] + float(r[iw].replace(',', '')) * UNIT[units[iw]]


def estep_bounds(path):
    """Per E-step kernel: executed warp instructions, duration, SM count x 4 schedulers, SM clock, tensor-pipe share."""
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {k: hdr.index(k) for k in ('Kernel Name', 'smsp__inst_executed.sum', 'gpu__time_duration.sum',
                                     'sm__cycles_elapsed.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
                                     'smsp__issue_active.avg.pct_of_peak_sustained_active')}
    out = {}
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('<unnamed>::', '').replace(' ', '')
        f = lambda k: float(r[col[k]].replace(',', ''))
        out[name] = {'warp_instructions': f('smsp__inst_executed.sum'), 'sm_cycles': f('sm__cycles_elapsed.avg'),
                     'duration_us': f('gpu__time_duration.sum'),
                     'tensor_pipe_active_pct': f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                     'issue_active_pct': f('smsp__issue_active.avg.pct_of_peak_sustained_active')}
    return out


def main(argv):
    tot, cnt = collections.defaultdict(float), collections.Counter()
    out = {}
    i = 0
    while i < len(argv):
        if argv[i] == '--estep':
            out['arhmm_estep'] = sum(b for _, b in launches(argv[i + 1]))
            out['arhmm_estep_bounds'] = estep_bounds(argv[i + 1])
            i += 2
            continue
        for name, b in launches(argv[i]):
            tot[name] += b
            cnt[name] += 1
        i += 1
    out.update({k: tot[k] / cnt[k] for k in sorted(tot)})
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main(sys.argv[1:])
