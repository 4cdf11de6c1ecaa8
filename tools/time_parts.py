"""Profiling aid: where does one bench step (2048-ROI bin) go?  preprocess vs network vs copies."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifcb_classifier_b200.engine import BinClassifier
from ifcb_classifier_b200 import preprocess as pp
from tests.fixtures import ref_model
import bench
dev = torch.device('cuda:0')
eng = BinClassifier('inception_v3', ref_model('inception_v3', 100, seed=0).state_dict(), device=dev, batch_cap=512, max_rois=2048)
b = bench.synth_bins(1, 2048)[0]
n, nb = eng.upload(b['roi'], b['offsets'], b['heights'], b['widths'])
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(3): eng.classify_device(n, nb)
torch.cuda.synchronize()
e = [ev() for _ in range(4)]
tp = tn = 0.0
reps = 5
for _ in range(reps):
    for i in range(0, n, 512):
        e[0].record()
        pp.preprocess_rois(eng.d_roi[:nb], eng.d_off[i:i+512], eng.d_h[i:i+512], eng.d_w[i:i+512], eng.R, out_mode=pp.OUT_U8_GRAY, out=eng.net.inp[:512])
        e[1].record()
        eng.net.forward(512)
        e[2].record()
        torch.cuda.synchronize()
        tp += e[0].elapsed_time(e[1]); tn += e[1].elapsed_time(e[2])
print('preprocess %.3f ms/batch, network %.3f ms/batch (512 ROIs)' % (tp / reps / 4, tn / reps / 4))
e0, e1 = ev(), ev()
e0.record()
for _ in range(reps): eng.classify_device(n, nb)
e1.record(); torch.cuda.synchronize()
print('classify_device %.3f ms/bin -> %.0f ROI/s' % (e0.elapsed_time(e1) / reps, n * reps / e0.elapsed_time(e1) * 1e3))
