"""Condensed view of an `ncu --page source --csv` dump: runs of SASS instructions with the same execution count.
    python tools/src_profile.py gpurun_out/src_X.csv [min_share_percent]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [(r[ia], r[isrc], int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex]
tot = sum(d[2] for d in data)
tots = sum(d[3] for d in data)
thresh = float(sys.argv[2]) / 100 if len(sys.argv) > 2 else 0.003
print('total inst', tot, 'samples', tots, 'n', len(data))
run, out = None, []
for i, (a, s, e, sm) in enumerate(data):
    if run and abs(e - run['e']) <= 0.02 * max(e, run['e'], 1):
        run['n'] += 1; run['tot'] += e; run['smp'] += sm; run['last'] = s
    else:
        if run:
            out.append(run)
        run = {'start': i, 'e': e, 'n': 1, 'tot': e, 'smp': sm, 'first': s, 'last': s}
out.append(run)
for r in out:
    if r['tot'] > tot * thresh or r['smp'] > tots * thresh:
        print('%5d n=%4d exec/inst=%11d  tot=%5.1f%%  smp=%5.1f%%  %s' % (
            r['start'], r['n'], r['e'], 100 * r['tot'] / tot, 100 * r['smp'] / tots, r['first'][:70]))
