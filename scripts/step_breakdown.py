#!/usr/bin/env python
"""Per-kernel breakdown of ONE training step from an ncu launch list (gpu__time_duration.sum CSV)."""
import collections, csv, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def us(x):
    t = float(x['Metric Value'].replace(',', '')); u = x['Metric Unit']
    return t / 1e3 if u == 'ns' else (t * 1e3 if u == 'ms' else t)
# one step = from one "first kernel of a step" to the launch before the next one
first = 'renorm_rows_kernel' if any('renorm_rows_kernel' in x['Kernel Name'] for x in rows) else 'catalog_prep_fwd'
starts = [i for i, x in enumerate(rows) if first in x['Kernel Name']]
s, e = starts[-2], starts[-1] - 1
agg = collections.OrderedDict(); tot = 0
detail = '-v' in sys.argv
for x in rows[s:e + 1]:
    n = x['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += us(x); tot += us(x)
    if detail: print(f"{n[:60]:62s} {x['Grid Size']:>16s} {us(x):9.1f}")
print(f'one step = {e - s + 1} launches, {tot:.0f} us summed (serialised, cold-cache)')
print('| kernel | launches | us | share |\n|---|---|---|---|')
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{n}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |')
