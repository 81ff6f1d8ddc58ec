"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the captured step)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg, tot = collections.OrderedDict(), 0.0
for r in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    v = float(r['Metric Value'].replace(',', ''))
    us = v / 1000 if r['Metric Unit'] in ('ns', 'nsecond') else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    tot += us
print('captured %d launches, %.1f us total (ncu: cold-cache, serialised -- compare shares, not absolutes)' % (sum(a[0] for a in agg.values()), tot))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-64s n=%4d  us=%9.1f  share=%5.1f%%' % (k[:64], n, us, 100 * us / tot))
