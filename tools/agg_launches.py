"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0     # launches to skip (build / warm-up)
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name']
    m = re.search(r'((?:ifcb|at)::(?:\(anonymous namespace\)::|native::)*\w+)', name)
    name = m.group(1).replace('(anonymous namespace)::', '') if m else name[:60]
    tm = re.search(r'<([^>]*)>', row['Kernel Name'])
    if tm and name.startswith('ifcb'):
        name += '<' + tm.group(1)[:12] + '>'
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    us = v / 1e3 if unit in ('ns', 'nsecond') else v if unit in ('us', 'usecond') else v * 1e3
    rows.append((int(row['ID']), name, us))
rows = rows[skip:]
agg = collections.OrderedDict()
for _, n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print('%d launches, %.3f ms total' % (len(rows), tot / 1e3))
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-50s %5d launches %10.3f ms %5.1f%%' % (n[:50], c, us / 1e3, 100 * us / tot))
