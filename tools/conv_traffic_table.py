"""Per-launch table of the conv kernels of one 1024-ROI Inception-v3 batch from the ncu CSV that tools/gpurun_jobs/r02_evidence2.sh
captures (gpu__time_duration, dram bytes read + written, tensor-pipe active %, L2 hit rate per launch), with the layer names of the
per-layer CUDA-event table.  Writes the .txt table and the .json that bench.py reads for `roofline.traffic`.

    python tools/conv_traffic_table.py gpurun_out/r02_conv_traffic.csv gpurun_out/r02_layer_events_final_b1024.txt profiles/r02
"""
import csv
import json
import sys
from collections import defaultdict

csv_path, layers_path, out_prefix = sys.argv[1:4]
names = [l.split()[0] for l in open(layers_path) if len(l.split()) > 2 and l.split()[1] == 'conv']
rows = defaultdict(dict)
for r in csv.DictReader(l for l in open(csv_path) if l.startswith('"')):
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    m = r['Metric Name']
    if m == 'gpu__time_duration.sum':
        v = v / 1e3 if u.startswith('n') else v if u.startswith('u') else v * 1e3
    elif m.startswith('dram__bytes'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    rows[int(r['ID'])][m] = v
ids = sorted(rows)
assert len(ids) == len(names), (len(ids), len(names))
total = 0.0
with open(out_prefix + '_ncu_conv_traffic.txt', 'w') as f:
    f.write('# per conv launch of one 1024-ROI Inception-v3 batch (fp16 operands): ncu duration, DRAM read+write bytes, tensor-pipe active, L2 hit rate\n')
    for i, n in zip(ids, names):
        r = rows[i]
        b = r['dram__bytes_read.sum'] + r['dram__bytes_write.sum']
        total += b
        f.write('%-32s %8.1f us  %8.1f MB  tensor-pipe active %5.1f%%  L2 hit %5.1f%%\n' %
                (n, r['gpu__time_duration.sum'], b / 1e6, r['sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'],
                 r['lts__t_sector_hit_rate.pct']))
json.dump({'batch': 1024, 'model': 'inception_v3', 'dtype': 'fp16', 'conv_launches': len(ids), 'dram_bytes_total': total,
           'dram_bytes_per_launch': total / len(ids), 'source': out_prefix + '_ncu_conv_traffic.csv'}, open(out_prefix + '_conv_traffic.json', 'w'))
print(len(ids), 'launches', '%.2f GB' % (total / 1e9))
