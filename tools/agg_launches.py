"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
agg = collections.OrderedDict()
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1.0)
    if len(sys.argv) > 2 and sys.argv[2] == 'grid':
        name += ' grid=' + d['Grid Size']
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-60s n=%5d  %10.3f ms  %5.1f%%  avg %9.1f us" % (k[:60], c, v / 1e6, 100 * v / tot, v / c / 1e3))
print("total ms %.3f" % (tot / 1e6))
