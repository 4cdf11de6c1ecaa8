"""Profiling helper: builds the Inception-v3 (or other) plan and runs it `--passes` times
(no bench logic), so ncu can capture individual layer launches by index.
    ncu --set full -k regex:conv_umma -s <skip> -c <n> python tools/run_plan_once.py --batch 512
Conv launches per Inception pass: 65 (order = graph order; see profiles/*layer_breakdown*)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifcb_classifier_b200.graph import CompiledNet  # noqa: E402
from tests.fixtures import ref_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--model', default='inception_v3')
ap.add_argument('--batch', type=int, default=512)
ap.add_argument('--passes', type=int, default=2)
ap.add_argument('--dtype', default='fp16')
ap.add_argument('--time', action='store_true', help='print per-layer CUDA-event times of the last pass')
a = ap.parse_args()
dev = torch.device('cuda:0')
net = CompiledNet(a.model, ref_model(a.model, 100).state_dict(), a.batch, in_kind='u8', device=dev, dtype=a.dtype)
net.inp.random_(0, 256)
for _ in range(a.passes):
    net.forward(a.batch)
torch.cuda.synchronize()
if a.time:
    pb = net.pb
    n = len(pb.layer_names)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    import numpy as np
    acc = np.zeros(n)
    for rep in range(4):
        ev[0].record()
        for i in range(n):
            pb.run(a.batch, i, i + 1)
            ev[i + 1].record()
        torch.cuda.synchronize()
        if rep:
            acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)])
    acc /= 3
    tot = acc.sum()
    for i in range(n):
        fl = pb.layer_flops[i] * a.batch
        print('%-34s %-5s %9.1f us %5.1f%% %8.1f TFLOP/s' % (pb.layer_names[i], pb.layer_kinds[i], acc[i] * 1e3,
                                                            100 * acc[i] / tot, fl / (acc[i] * 1e-3) / 1e12 if acc[i] > 0 else 0))
    print('TOTAL %.1f us  -> %.0f ROI/s' % (tot * 1e3, a.batch / (tot * 1e-3)))
