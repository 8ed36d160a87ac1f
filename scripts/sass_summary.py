"""Per-kernel counts of the Blackwell-specific SASS instructions in libbehavenet_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (cp.async.bulk.tensor), UBLKCP, UTCBAR
(tcgen05.commit), SYNCS (mbarrier), LDGSTS (cp.async), plus one excerpt of an MMA issue sequence.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'behavenet_b200', 'libbehavenet_b200.so')
PAT = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'LDGSTS', 'HMMA', 'FFMA2', 'FFMA', 'DFMA',
       'REDG', 'ATOMG', 'RED.', 'UTCATOMSWS']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    fn = None
    counts = collections.OrderedDict()
    excerpt = []
    want = None
    for line in out.split('\n'):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r'\(.*', '', fn.replace('(anonymous namespace)::', '').replace('void ', ''))
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        ins = re.search(r'/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if not ins:
            continue
        op = ins.group(2)
        counts[fn]['total'] += 1
        for p in PAT:
            if op.startswith(p):
                counts[fn][p.rstrip('.')] += 1
        if 'igemm_tma_kernel<128, 2, 2>' in fn and op.startswith('UTCHMMA') and want is None:
            want = 14
        if want:
            excerpt.append(line.rstrip())
            want -= 1
    print('libbehavenet_b200.so, cuobjdump -sass: instruction counts per kernel (sm_100a)\n')
    cols = ['total', 'UTCHMMA', 'LDTM', 'UTMALDG', 'UTCBAR', 'SYNCS', 'LDGSTS', 'FFMA2', 'FFMA', 'DFMA', 'HMMA']
    print('%-62s' % 'kernel' + ''.join('%9s' % c for c in cols))
    for fn, c in counts.items():
        if c['total'] == 0:
            continue
        print('%-62s' % fn[:62] + ''.join('%9d' % c[k] for k in cols))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print('%-62s' % 'ALL KERNELS' + ''.join('%9d' % tot[k] for k in cols))
    print('\nUTCHMMA = tcgen05.mma (kind::tf32), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit,')
    print('SYNCS = mbarrier ops, LDGSTS = cp.async, FFMA2 = packed fp32 FMA; no HMMA (legacy mma.sync) anywhere.\n')
    print('excerpt: first tcgen05.mma issue of igemm_tma_kernel<128, 2, 2> (two accumulators per weight tile)')
    print('\n'.join(excerpt))


if __name__ == '__main__':
    main()
