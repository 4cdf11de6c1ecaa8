"""File-to-file RUN throughput: `neuston_net RUN` over synthetic bins on disk (wall clock, host I/O + result files included).
    python tools/bench_cli_run.py --bins 16 --outfile '{BIN_ID}_class.mat'"""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument('--bins', type=int, default=16)
ap.add_argument('--rois', type=int, default=2048)
ap.add_argument('--model', default='inception_v3')
ap.add_argument('--outfile', default='{BIN_ID}_class.mat')
ap.add_argument('--batch', type=int, default=1024)
a = ap.parse_args()
from oracle import synth_bins
from ifcb_classifier_b200 import neuston_net
from ifcb_classifier_b200.neuston_models import NeustonModel

tmp = tempfile.mkdtemp()
for i in range(a.bins):
    synth_bins.write_bin(os.path.join(tmp, 'bins'), synth_bins.make_bin(i, n_rois=a.rois))
torch.manual_seed(0)
hp = argparse.Namespace(MODEL=a.model, classes=['class_%03d' % i for i in range(100)], pretrained=False, resize=299 if a.model == 'inception_v3' else 224,
                        img_norm=None, model_id='bench', seed=1)
ckpt = os.path.join(tmp, 'bench.ptl')
NeustonModel(hp).save_checkpoint(ckpt)
argv = ['--batch', str(a.batch), 'RUN', os.path.join(tmp, 'bins'), ckpt, 'R', '--outdir', os.path.join(tmp, 'out'), '--outfile', a.outfile]
neuston_net.main(argv + ['--clobber'])                 # warm-up: plan build, first launches, page cache
t0 = time.perf_counter()
parser = neuston_net.argparse_nn()
args = parser.parse_args(argv + ['--clobber'])
clf = NeustonModel.load_from_checkpoint(args.MODEL)
neuston_net.argparse_nn_runtimeparams(args, clf)
t1 = time.perf_counter()
summ = neuston_net.do_run(args, clf)
t2 = time.perf_counter()
n = sum(s['n_rois'] for s in summ)
print('CLI_RUN %s %s: %d bins, %d ROIs; do_run %.2f s -> %.0f ROI/s file to file (plan build included), of which classify loop %.2f s -> %.0f ROI/s'
      % (a.model, a.outfile, a.bins, n, t2 - t1, n / (t2 - t1), summ[0]['seconds'], n / summ[0]['seconds']))
