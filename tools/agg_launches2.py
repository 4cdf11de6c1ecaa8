"""Per-launch view of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` list:
the last `--step-launches` launches (one training step), aggregated by kernel and listed individually by time."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = defaultdict(dict)
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    k = int(r['ID'])
    rows[k]['name'] = r['Kernel Name']
    rows[k]['grid'] = r['Grid Size']
    rows[k][r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    rows[k]['unit_' + r['Metric Name']] = r['Metric Unit']
ids = sorted(rows)
# a step starts at the last launch of stem_im2col_kernel / the memset before it: keep the launches from the LAST stem_im2col on
start = max(i for i in ids if 'stem_im2col' in rows[i]['name'])
step = [rows[i] for i in ids if i >= start - 2]
def us(r):
    v = r['gpu__time_duration.sum']
    u = r['unit_gpu__time_duration.sum']
    return v / 1e3 if u.startswith('n') else v if u.startswith('u') else v * 1e3
def short(n):
    n = re.sub(r'ifcb::|<unnamed>::|\(anonymous namespace\)::|void ', '', n)
    return re.sub(r'\(.*', '', n)[:46]
tot = sum(us(r) for r in step)
print('%d launches in the step, %.2f ms of kernel time (cold-cache, serialised)' % (len(step), tot / 1e3))
agg = defaultdict(lambda: [0, 0.0, 0.0])
for r in step:
    a = agg[short(r['name'])]
    a[0] += 1; a[1] += us(r); a[2] += r.get('dram__bytes_read.sum', 0) + r.get('dram__bytes_write.sum', 0)
for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-46s %4d launches %9.1f us %5.1f%%  %7.2f GB  %5.2f TB/s' % (n, c, t, 100 * t / tot, b / 1e9, b / 1e12 / (t * 1e-6) if t else 0))
print('--- top launches ---')
for r in sorted(step, key=lambda r: -us(r))[:topn]:
    b = r.get('dram__bytes_read.sum', 0) + r.get('dram__bytes_write.sum', 0)
    print('%-46s grid %-14s %8.1f us  %7.1f MB  %5.2f TB/s' % (short(r['name']), r['grid'], us(r), b / 1e6, b / 1e12 / (us(r) * 1e-6)))
