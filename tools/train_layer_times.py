"""Per-launch CUDA-event times of one TRAIN step (profiling aid): every forward / backward closure of a TrainNet is timed on its own
(eager, warm L2), labelled by what it calls; weight-gradient launches carry their shape and the time their MMAs alone would take
(pixel tiles x block pairs x output-channel tiles x 8 K steps x max(N/2, 32 + N/4) clk over 148 SMs).

    python tools/train_layer_times.py --arch inception_v3 --batch 256
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifcb_classifier_b200 import _lib  # noqa: E402
from ifcb_classifier_b200.train import TrainNet  # noqa: E402
from tests.fixtures import ref_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--arch', default='inception_v3')
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--mhz', type=float, default=1900.0)
a = ap.parse_args()
dev = torch.device('cuda:0')
R = 299 if a.arch == 'inception_v3' else 224
net = TrainNet(a.arch, ref_model(a.arch, 100, seed=0).state_dict(), a.batch, device=dev, dtype='bf16', R=R)
g = torch.Generator().manual_seed(0)
x = torch.rand(a.batch, 3, R, R, generator=g).to(dev)
y = torch.randint(0, 100, (a.batch,), generator=g).to(dev)
for _ in range(2):
    net.forward_backward(x, y)
torch.cuda.synchronize()


def describe(fn):
    name = getattr(fn, '__name__', '?')
    for c in (fn.__closure__ or ()):
        v = c.cell_contents
        if isinstance(v, _lib.WgradDesc):
            return 'wgrad', v
    return name, None


def wgrad_model_us(d):
    P = (d.H + 2 * d.pad_h - d.kh) // d.stride_h + 1
    Q = (d.W + 2 * d.pad_w - d.kw) // d.stride_w + 1
    tiles = (d.batch * P * Q + 127) // 128
    pairs = (d.kh * d.kw * ((d.Cin + 63) // 64) + 1) // 2
    c16 = (d.Cout + 15) & ~15
    t = (c16 + 255) // 256
    tn = (((c16 + t - 1) // t) + 15) & ~15
    cot = (d.Cout + tn - 1) // tn
    clk = tiles * pairs * cot * 8 * max(tn / 2, 32 + tn / 4) / 148
    return clk / a.mhz, 2.0 * d.batch * P * Q * d.Cout * d.kh * d.kw * d.Cin


rows = []
for phase, fns in (('fwd', net.fwd), ('bwd', net.bwd)):
    for i, fn in enumerate(fns):
        kind, wd = describe(fn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        rows.append((phase, i, kind, e0.elapsed_time(e1) / a.reps * 1e3, wd))
tot = sum(r[3] for r in rows)
print('%s batch %d: %d closures, %.2f ms summed (eager, each timed alone, warm L2)' % (a.arch, a.batch, len(rows), tot / 1e3))
by = {}
for r in rows:
    by.setdefault((r[0], r[2]), [0, 0.0])
    by[(r[0], r[2])][0] += 1
    by[(r[0], r[2])][1] += r[3]
for k, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print('  %-4s %-14s %4d calls %9.1f us %5.1f%%' % (k[0], k[1], n, us, 100 * us / tot))
print('weight-gradient launches (backward order):')
for r in rows:
    if r[4] is not None:
        d = r[4]
        mu, fl = wgrad_model_us(d)
        print('  #%-4d %4dx%-4d %dx%d s%d  %4d -> %-4d  %8.1f us  %6.1f TFLOP/s  MMA-bound %7.1f us  x%.2f' %
              (r[1], d.H, d.W, d.kh, d.kw, d.stride_h, d.Cin, d.Cout, r[3], fl / r[3] / 1e6, mu, r[3] / max(mu, 1e-9)))
