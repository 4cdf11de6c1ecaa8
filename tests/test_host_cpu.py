"""CPU: host logic -- bin ingest vs the oracle's pyifcb stand-in, sharding (incl. a
world_size-2 gloo run), result files, checkpoint round trip, CLI surface."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from ifcb_classifier_b200 import ifcb_io, neuston_net, results, sharding
from oracle import ifcb_stub, synth_bins

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rawbin_matches_oracle_bin_reader(tmp_path):
    for idx, kw in ((0, dict(n_rois=40, empty_every=6)), (1, dict(n_rois=1)), (2, dict(n_rois=0))):
        synth_bins.write_bin(str(tmp_path / 'a' / 'b'), synth_bins.make_bin(idx, **kw))
    got = list(ifcb_io.DataDirectory(str(tmp_path)))
    want = list(ifcb_stub.DataDirectory(str(tmp_path)))
    assert [str(b.pid) for b in got] == [str(b.pid) for b in want] and len(got) == 3
    for g, w in zip(got, want):
        assert g.pids == [w.pid.with_target(t) for t in w.images]
        for i, (t, im) in enumerate(w.images.items()):
            assert np.array_equal(g.image(i), im)
        assert (g.pid.year, g.pid.yearday) == (w.pid.year, w.pid.yearday)
    assert len(got[2]) == 0      # empty bin: no ROIs, still enumerated (reported as 'Bin is Empty' by RUN)


def test_adc_table_outside_roi_is_rejected(tmp_path):
    b = synth_bins.make_bin(4, n_rois=5)
    b['roi'] = b['roi'][:-10]
    base = synth_bins.write_bin(str(tmp_path), b)
    with pytest.raises(ValueError):
        ifcb_io.RawBin(base)


def test_lpt_sharding_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    costs = [int(c) for c in rng.integers(1, 4000, 101)]
    for world in (1, 2, 4, 8):
        owner = sharding.assign(costs, world)
        assert sorted(set(owner)) == list(range(world)) and len(owner) == len(costs)
        load = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(world)]
        assert max(load) - min(load) <= max(costs)                # LPT bound
    assert sharding.assign([5, 5, 5, 5], 2) == [0, 1, 0, 1]           # deterministic tie-break


_GLOO_WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from ifcb_classifier_b200 import sharding, ifcb_io
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%(port)d', rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
bases = ifcb_io.DataDirectory(%(src)r).basepaths()
mine = sharding.my_bins(bases, rank, 2)
summ = sharding.gather_summary(dict(rank=rank, n_bins=len(mine), n_rois=sum(len(ifcb_io.RawBin(b)) for b in mine),
                                    seconds=0.0, error_bins=[], mine=[os.path.basename(b) for b in mine]), 2)
if rank == 0:
    print('SUMMARY ' + json.dumps(summ))
dist.destroy_process_group()
'''


def test_two_rank_gloo_bin_sharding(tmp_path):
    sizes = [30, 5, 17, 9, 22, 1, 14]
    for i, n in enumerate(sizes):
        synth_bins.write_bin(str(tmp_path / 'bins'), synth_bins.make_bin(i, n_rois=n))
    code = _GLOO_WORKER % dict(root=ROOT, port=29500 + os.getpid() % 2000, src=str(tmp_path / 'bins'))
    procs = [subprocess.Popen([sys.executable, '-c', code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith('SUMMARY ')][0]
    summ = json.loads(line[len('SUMMARY '):])
    all_bins = sorted(summ[0]['mine'] + summ[1]['mine'])
    assert all_bins == sorted(synth_bins.bin_lid(i) for i in range(len(sizes)))     # disjoint + complete
    assert summ[0]['n_rois'] + summ[1]['n_rois'] == sum(sizes)
    assert abs(summ[0]['n_rois'] - summ[1]['n_rois']) <= max(sizes)


def test_result_files_json_and_mat(tmp_path):
    from scipy.io import loadmat
    pid = ifcb_io.Pid('D20260101T000000_IFCB999'); pid.namespace = 'sub/'
    rng = np.random.default_rng(0)
    scores = rng.random((5, 3)).astype(np.float32); scores /= scores.sum(1, keepdims=True)
    imgs = [pid.with_target(t) for t in (1, 2, 4, 7, 9)]
    labels = ['a', 'b', 'c']
    p = results.save_run_results(imgs, scores, labels, '2026-01-01T00:00:00+00:00', str(tmp_path),
                                 'D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.json', 'm1', pid)
    assert p.endswith('D2026/D20260101/D20260101T000000_IFCB999_class.json')
    j = json.load(open(p))
    assert j['version'] == 'v3' and j['roi_numbers'] == [1, 2, 4, 7, 9] and j['bin_id'] == pid.pid
    assert j['output_classes'] == scores.argmax(1).tolist() and j['class_labels'] == labels
    p = results.save_run_results(imgs, scores, labels, 'ts', str(tmp_path), '{INPUT_SUBDIRS}/{BIN_ID}.mat', 'm1', pid)
    m = loadmat(p)
    assert (m['output_classes'].ravel() == scores.argmax(1) + 1).all()           # 1-based for matlab
    assert m['output_scores'].dtype == np.float32 and m['roi_numbers'].ravel().tolist() == [1, 2, 4, 7, 9]
    with pytest.raises(AssertionError):
        results.save_run_results(imgs, scores, labels + ['d'], 'ts', str(tmp_path), '{BIN_ID}.json', 'm1', pid)
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError):
            results.save_run_results(imgs, scores, labels, 'ts', str(tmp_path), '{BIN_ID}_class.h5', 'm1', pid)


def test_model_names_and_checkpoint_roundtrip(tmp_path):
    from ifcb_classifier_b200.neuston_models import NeustonModel, get_namebrand_model
    with pytest.raises(KeyError, match='model unknown!'):
        get_namebrand_model('nope', 3)
    torch.manual_seed(0)
    hp = dict(MODEL='resnet18', classes=['x', 'y', 'z'], pretrained=False, resize=224, img_norm=['0.667', '0.161'],
              model_id='unit', seed=0)
    m = NeustonModel(hp)
    assert m.model.state_dict()['fc.weight'].shape == (3, 512) and not m.model.transform_input
    path = str(tmp_path / 'unit.ptl')
    m.save_checkpoint(path)
    ck = torch.load(path, weights_only=False)
    assert all(k.startswith('model.') for k in ck['state_dict']) and ck['hyper_parameters']['model_id'] == 'unit'
    m2 = NeustonModel.load_from_checkpoint(path)
    assert m2.hparams.classes == ['x', 'y', 'z'] and m2.hparams.img_norm == ['0.667', '0.161']
    for k, v in m.model.state_dict().items():
        assert torch.equal(v, m2.model.state_dict()[k])
    inc = get_namebrand_model('inception_v3', 5, False)
    assert inc.state_dict()['AuxLogits.fc.weight'].shape[0] == 5 and inc.state_dict()['fc.weight'].shape == (5, 2048)


def test_cli_surface_matches_reference_flags():
    p = neuston_net.argparse_nn()
    a = p.parse_args(['--batch', '64', 'RUN', 'src', 'model.ptl', 'rid', '--filter', 'IN', 'D2026', '--clobber',
                      '--outfile', '{BIN_ID}.json', '--outfile', '{BIN_ID}.mat'])
    assert (a.cmd_mode, a.batch_size, a.loaders, a.src_type) == ('RUN', 64, 4, 'bin')
    assert a.outdir == 'run-output/{RUN_ID}/v3/{MODEL_ID}' and a.outfile == ['{BIN_ID}.json', '{BIN_ID}.mat']
    t = p.parse_args(['TRAIN', 'src', 'inception_v3', 'tid', '--untrain', '--img-norm', '0.667', '0.161', '--flip', 'xy+V'])
    assert (t.batch_size, t.pretrained, t.emax, t.emin, t.estop, t.split, t.class_min) == (108, False, 60, 10, 10, '80:20', 2)
    assert t.outdir == 'training-output/{TRAIN_ID}' and t.model_id == '{TRAIN_ID}' and t.epochs_log == 'epochs.csv'


# ---- TRAIN: gradient buckets + data-parallel mean (world_size-2 gloo, CPU) ----
def test_plan_buckets_tile_the_arena():
    # offsets of the backward units' first parameter, as TrainNet lays them out (forward order => decreasing)
    offs = [900, None, 700, None, None, 650, 300, None, 0]
    marks = sharding.plan_buckets(offs, 1000, 250)
    assert marks == [(2, 700, 1000), (6, 300, 700), (8, 0, 300)]
    covered = sorted((lo, hi) for _, lo, hi in marks)
    assert covered[0][0] == 0 and covered[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    assert sharding.plan_buckets(offs, 1000, 10 ** 9) == [(8, 0, 1000)]          # one bucket when the threshold is never met
    assert sharding.plan_buckets([None, None], 0, 4) == []


_GLOO_TRAIN_WORKER = r'''
import sys, json
sys.path.insert(0, %(root)r)
import torch
import torch.distributed as dist
from ifcb_classifier_b200 import sharding
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%(port)d', rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
n = 1000
g = torch.Generator().manual_seed(100 + rank)
grads = torch.randn(n, generator=g)
mine = grads.clone()
red = sharding.GradReducer(grads)
offs = [900, None, 700, None, None, 650, 300, None, 0]
marks = sharding.plan_buckets(offs, n, 250)
by_unit = {}
for i, lo, hi in marks:
    by_unit.setdefault(i, []).append((lo, hi))
for unit in range(len(offs)):                # the backward pass: a bucket is reduced as soon as its last unit ran
    for lo, hi in by_unit.get(unit, ()):
        red(lo, hi)
scale = red.wait()
other = torch.randn(n, generator=torch.Generator().manual_seed(100 + (1 - rank)))
want = (mine + other) * 0.5
ok = bool(torch.allclose(grads * scale, want, rtol=0, atol=1e-6)) and scale == 0.5 and red.world == 2 and not red.works
print('RESULT ' + json.dumps(dict(rank=rank, ok=ok)))
dist.destroy_process_group()
'''


def test_two_rank_gloo_gradient_mean():
    code = _GLOO_TRAIN_WORKER % dict(root=ROOT, port=31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, '-c', code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    for out, _ in outs:
        line = [l for l in out.splitlines() if l.startswith('RESULT ')][0]
        assert json.loads(line[len('RESULT '):])['ok']


def test_result_files_for_image_inputs(tmp_path):
    """RUN --type img (neuston_callbacks.py:186-206): one file with input_images, or one per input sub-directory."""
    rng = np.random.default_rng(1)
    scores = rng.random((4, 3)).astype(np.float32)
    src = str(tmp_path / 'imgs') + os.sep
    imgs = [src + 'a/x1.png', src + 'a/x2.png', src + 'b/c/y1.png', src + 'b/c/y2.png']
    for p_ in imgs:
        os.makedirs(os.path.dirname(p_), exist_ok=True)
        open(p_, 'wb').close()
    p = results.save_run_results(imgs, scores, ['u', 'v', 'w'], 'ts', str(tmp_path / 'out'), 'img_results.json', 'm1', src)
    j = json.load(open(p))
    assert j['input_images'] == imgs and 'bin_id' not in j and j['output_classes'] == scores.argmax(1).tolist()
    ps = results.save_run_results(imgs, scores, ['u', 'v', 'w'], 'ts', str(tmp_path / 'out2'), '{INPUT_SUBDIRS}/res.json', 'm1', src)
    assert sorted(os.path.relpath(q, str(tmp_path / 'out2')) for q in ps) == ['a/res.json', 'b/c/res.json']
    j = json.load(open([q for q in ps if q.endswith('b/c/res.json')][0]))
    assert j['input_images'] == ['y1.png', 'y2.png'] and np.allclose(j['output_scores'], scores[2:])
