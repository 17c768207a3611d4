"""Top stall sites of each kernel in an `ncu --page source --csv` export.  Usage: ncu_top_stalls.py sass.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1][:70], 'rows': []}
        kern.append(cur)
    elif r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and len(r) > 5:
        cur['rows'].append(r)
for k in kern:
    h = k['hdr']
    si, so, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    tot = sum(int(r[si]) for r in k['rows'])
    print('=====', k['name'], 'samples', tot, 'instr', len(k['rows']))
    stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    agg = {}
    for r in k['rows']:
        for i in stall_cols:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print('  totals:', sorted(((v, c) for c, v in agg.items()), reverse=True)[:8])
    for idx, r in sorted(enumerate(k['rows']), key=lambda t: -int(t[1][si]))[:n]:
        st = sorted([(int(r[i]), h[i][6:]) for i in stall_cols], reverse=True)[:2]
        print(f"{idx:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}% ex={r[ie]:>8s} {r[so].strip()[:64]:64s} {st}")
