"""Hot SASS lines of an ncu report's source page: python scripts/ncu_hot.py report.ncu-rep [N] [kernel-substr]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for blk in blocks[1:]:
    lines = blk.split('\n')
    name = lines[0]
    if len(sys.argv) > 3 and sys.argv[3] not in name:
        continue
    rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
    hdr = rows[0]
    si = hdr.index('# Samples'); ie = hdr.index('Instructions Executed'); so = hdr.index('Source')
    body = [r for r in rows[1:] if len(r) > si and r[si] not in ('',)]
    tot = sum(float(r[si]) for r in body) or 1
    print('==', name[:100], 'samples', tot, 'sass lines', len(body))
    idx = {id(r): i for i, r in enumerate(body)}
    for r in sorted(body, key=lambda r: -float(r[si]))[:N]:
        print('%5d %7.0f %5.1f%% %10s  %s' % (idx[id(r)], float(r[si]), 100 * float(r[si]) / tot, r[ie], r[so].strip()[:100]))
