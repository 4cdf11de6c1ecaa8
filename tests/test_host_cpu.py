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


def test_result_file_h5_default_layout(tmp_path):
    """The reference's DEFAULT output (neuston_net.py:180-182, _save_run_results_hdf neuston_callbacks.py:252-268): an empty
    float32 'metadata' dataset with version / model_id / timestamp / bin_id attributes, float16 gzip output_classes and
    output_scores, variable-length-string class_labels, uint16 roi_numbers -- written without h5py and parsed back by the
    independent reader (superblock, symbol table, object headers, chunk B-trees, global heap, deflate)."""
    from ifcb_classifier_b200 import h5lite
    pid = ifcb_io.Pid('D20260101T000000_IFCB999'); pid.namespace = ''
    rng = np.random.default_rng(1)
    scores = rng.random((2048, 100)).astype(np.float32); scores /= scores.sum(1, keepdims=True)
    targets = np.arange(1, 2049) + (np.arange(2048) // 7)
    imgs = [pid.with_target(t) for t in targets]
    labels = ['class_%03d' % i for i in range(99)] + ['Dinobryon_\u00e9']
    p = results.save_run_results(imgs, scores, labels, '2026-01-01T00:00:00+00:00', str(tmp_path),
                                 'D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.h5', 'm1', pid)
    assert p.endswith('D2026/D20260101/D20260101T000000_IFCB999_class.h5')
    with open(p, 'rb') as f:
        assert f.read(8) == b'\x89HDF\r\n\x1a\n'
    f = h5lite.read(p)
    assert sorted(f) == ['class_labels', 'metadata', 'output_classes', 'output_scores', 'roi_numbers']
    assert f['metadata'].shape is None and f['metadata'].dtype == np.float32
    assert f['metadata'].attrs == dict(version='v3', model_id='m1', timestamp='2026-01-01T00:00:00+00:00', bin_id=pid.pid)
    assert f['output_scores'].dtype == np.float16 and np.array_equal(f['output_scores'].data, scores.astype(np.float16))
    assert f['output_classes'].dtype == np.float16 and np.array_equal(f['output_classes'].data, scores.argmax(1).astype(np.float16))
    assert f['roi_numbers'].dtype == np.uint16 and f['roi_numbers'].data.tolist() == targets.tolist()
    assert f['class_labels'].data.tolist() == labels
    assert all(d.filters == [(1, (4,))] for k, d in f.items() if k != 'metadata')            # deflate on every array
    # --type img files carry input_images instead of bin_id / roi_numbers
    p2 = results.save_run_results(['/x/a.png', '/x/b.png'], scores[:2], labels, 'ts', str(tmp_path), 'img_results.h5', None, '/x')
    g = h5lite.read(p2)
    assert g['input_images'].data.tolist() == ['/x/a.png', '/x/b.png'] and 'bin_id' not in g['metadata'].attrs
    assert g['metadata'].attrs['model_id'] == ''


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


# ---- TRAIN host logic (neuston_data mirrors), CPU only ----
def _png_tree(root, counts):
    from PIL import Image
    for cls, n in counts.items():
        d = os.path.join(root, cls)
        os.makedirs(d, exist_ok=True)
        for i in range(n):
            Image.fromarray(np.full((5 + i % 3, 7), i, np.uint8), mode='L').save(os.path.join(d, '%s_%03d.png' % (cls, i)))
        open(os.path.join(d, 'notes.txt'), 'w').close()              # non-image files are ignored


def test_neuston_dataset_thresholds_and_split(tmp_path):
    """NeustonDataset (reference neuston_data.py:20-187): class folders, --class-min / --class-max, seeded T:V split."""
    import random
    import argparse
    from ifcb_classifier_b200 import neuston_data as nd
    src = str(tmp_path / 'ds')
    _png_tree(src, dict(big=20, mid=10, tiny=1))
    full = nd.NeustonDataset(src, minimum_images_per_class=2)
    assert full.classes == ['big', 'mid'] and full.classes_ignored_from_too_few_samples == [('tiny', 1)]
    assert full.count_perclass == [20, 10] and len(full) == 30 and full.imgs is full.images
    random.seed(0)
    capped = nd.NeustonDataset(src, minimum_images_per_class=2, maximum_images_per_class=12)
    assert capped.count_perclass == [12, 10] and capped.classes_limited_from_too_many_samples == ['big']
    a, b = full.split(80, 20, seed=5)
    assert a.count_perclass == [16, 8] and b.count_perclass == [4, 2] and a.classes == b.classes
    assert not set(a.images) & set(b.images) and len(a) + len(b) == len(full)
    a2, _ = full.split(80, 20, seed=5)
    assert a2.images == a.images                                        # same seed, same split
    with pytest.raises(AssertionError):
        full.split(70, 20)
    img, target, path = full[0]
    assert img.dtype == np.uint8 and img.ndim == 2 and target == 0 and path.endswith('.png')
    # get_trainval_datasets: the reference's flag semantics ('x' -> vertical flip, 'y' -> horizontal, '+V' also on validation)
    args = argparse.Namespace(SRC=src, class_config=None, class_min=2, class_max=None, split='80:20', seed=5, swap=False,
                              MODEL='inception_v3', img_norm=['0.5', '0.25'], flip='x+V')
    tr, va = nd.get_trainval_datasets(args)
    assert args.resize == 299 and tr.transforms['flips'] == ['v'] and va.transforms['flips'] == ['v']
    assert tr.transforms['img_norm'] == ([0.5] * 3, [0.25] * 3) and len(tr) == 24 and len(va) == 6
    args.flip, args.MODEL, args.swap = 'xy', 'resnet18', True
    tr, va = nd.get_trainval_datasets(args)
    assert args.resize == 224 and tr.transforms['flips'] == ['v', 'h'] and va.transforms['flips'] == [] and len(tr) == 6


def test_neuston_dataset_class_config_csv(tmp_path):
    """--class-config CSV COL (reference neuston_data.py:189-256): 1 keep, 0 drop, other values rename / group."""
    from ifcb_classifier_b200 import neuston_data as nd
    src = str(tmp_path / 'ds')
    _png_tree(src, dict(a=4, b=3, c=5, d=2))
    cfg = tmp_path / 'cfg.csv'
    cfg.write_text('class,run1,run2\na,1,1\nb,0,grp\nc,grp,grp\nd,1,0\nmissing,1,1\n')
    ds = nd.NeustonDataset.from_csv(src, str(cfg), 'run1')
    assert ds.classes == ['a', 'd', 'grp'] and ds.count_perclass == [4, 2, 5]
    ds = nd.NeustonDataset.from_csv(src, str(cfg), 'run2', minimum_images_per_class=5)
    assert ds.classes == ['grp'] and ds.count_perclass == [8] and ds.classes_ignored_from_too_few_samples == [('a', 4)]


def test_image_batcher_indices_are_a_distributed_partition():
    """train_loop.ImageBatcher: shuffled per epoch, padded to a multiple of the world size and strided by rank."""
    from ifcb_classifier_b200.train_loop import ImageBatcher

    class DS(object):
        images, targets, transforms = list(range(10)), [0] * 10, None

        def __len__(self):
            return 10

    parts = []
    for r in range(4):
        b = ImageBatcher(DS(), 2, 'cpu', loaders=1, rank=r, world=4, shuffle=True, seed=3)
        parts.append(b.indices())
        assert len(b) == 2 and len(parts[-1]) == 3
    flat = [i for p_ in parts for i in p_]
    assert sorted(set(flat)) == list(range(10)) and len(flat) == 12            # complete; 2 wrapped-around items
    b0 = ImageBatcher(DS(), 2, 'cpu', loaders=1, shuffle=True, seed=3)
    e0 = b0.indices(); b0.epoch = 1
    assert sorted(e0) == list(range(10)) and b0.indices() != e0                 # reshuffled every epoch


def test_validation_results_files(tmp_path):
    """train_loop.save_validation_results (reference SaveValidationResults, neuston_callbacks.py:20-156): chosen series only,
    1-based list-typed class indices in .mat, confusion matrix / F1 from the validation scores."""
    from scipy.io import loadmat
    from ifcb_classifier_b200.train_loop import save_validation_results

    class DS(object):
        def __init__(self, images, targets):
            self.images, self.targets = images, targets
            self.count_perclass = [targets.count(0), targets.count(1), targets.count(2)]

    train = DS(['/d/a/t%d.png' % i for i in range(6)], [0, 0, 1, 1, 2, 2])
    val = DS(['/d/a/v%d.png' % i for i in range(4)], [0, 1, 2, 2])
    args = argparse.Namespace(classes=['a', 'b', 'c'], model_id='m', cmd_timestamp='ts', outdir=str(tmp_path))
    scores = np.array([[.8, .1, .1], [.2, .7, .1], [.1, .2, .7], [.6, .3, .1]], np.float32)     # last one is wrong (2 -> 0)
    series = 'training_image_basenames training_classes image_basenames input_classes output_scores confusion_matrix counts_perclass f1_perclass f1_weighted f1_macro'.split()
    p = save_validation_results('results.mat', series, args, 3, train, val, np.array([0, 1, 2, 2]), scores, val.images)
    m = loadmat(p)
    # upstream stores ndarray series as float32 unshifted (neuston_callbacks.py:130); only list-typed index series are 1-based
    assert m['input_classes'].ravel().tolist() == [0, 1, 2, 2] and m['output_classes'].ravel().tolist() == [0, 1, 2, 0]
    assert m['confusion_matrix'].tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 1]]
    assert m['counts_perclass'].ravel().tolist() == [3, 3, 4] and m['training_classes'].ravel().tolist() == [1, 1, 2, 2, 3, 3]
    assert [str(s[0]) for s in m['image_basenames'].ravel()] == ['v0', 'v1', 'v2', 'v3']
    assert abs(float(m['f1_macro'].ravel()[0]) - (2 / 3 + 1 + 2 / 3) / 3) < 1e-6 and 'recall_macro' not in m
    p = save_validation_results('e{epoch}/results.json', ['output_winscores', 'classes_by_count'], args, 3, train, val, np.array([0, 1, 2, 2]),
                                scores, val.images)
    j = json.load(open(p))
    assert p.endswith('e3/results.json') and j['classes_by_count'] == [2, 0, 1] and np.allclose(j['output_winscores'], [.8, .7, .7, .6])
    assert 'confusion_matrix' not in j and j['class_labels'] == ['a', 'b', 'c']
    from ifcb_classifier_b200 import h5lite
    p = save_validation_results('results.h5', series + ['classes_by_f1'], args, 0, train, val, np.array([0, 1, 2, 2]), scores, val.images)
    h = h5lite.read(p)
    assert h['input_classes'].dtype == np.int16 and h['input_classes'].data.tolist() == [0, 1, 2, 2]       # 0-based in .h5
    assert h['confusion_matrix'].dtype == np.float16 and h['confusion_matrix'].data.tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 1]]
    assert h['image_basenames'].data.tolist() == ['v0', 'v1', 'v2', 'v3'] and h['class_labels'].data.tolist() == ['a', 'b', 'c']
    assert abs(h['metadata'].attrs['f1_macro'] - (2 / 3 + 1 + 2 / 3) / 3) < 1e-9 and h['metadata'].attrs['model_id'] == 'm'
    assert h['f1_perclass'].dtype == np.float16 and h['counts_perclass'].data.tolist() == [3, 3, 4]


def test_json_score_text_is_byte_identical_to_python(built_lib):
    """ifcb_format_scores_json == json.dumps(scores.tolist()) byte for byte (float32 widened to float64, shortest repr,
    Python's fixed / exponent switch), so .json result files are what the reference writes (neuston_callbacks.py:213-230)."""
    rng = np.random.default_rng(0)
    cases = [rng.random((64, 100)).astype(np.float32), (10.0 ** rng.uniform(-45, 38, (200, 7))).astype(np.float32),
             np.array([[0.0, 1.0, 0.5, 1e-4, 9.99e-5, 1e-5, 123456.0, 1e15, 1e16, 1.5e16, 3.4028235e38, 1e-45, -0.0, -2.5, 0.1]], np.float32),
             -np.exp(rng.normal(0, 8, (50, 11))).astype(np.float32), np.zeros((0, 5), np.float32), np.zeros((3, 0), np.float32)]
    for a in cases:
        assert results._scores_json(a) == json.dumps(a.tolist())
    # whole file: same bytes as the plain encoder would produce
    pid = ifcb_io.Pid('D20260101T000000_IFCB999'); pid.namespace = ''
    scores = rng.random((9, 4)).astype(np.float32)
    imgs = [pid.with_target(t + 1) for t in range(9)]
    import tempfile
    d = tempfile.mkdtemp()
    p = results.save_run_results(imgs, scores, list('abcd'), 'ts', d, '{BIN_ID}.json', 'm', pid)
    want = json.dumps(dict(version='v3', model_id='m', timestamp='ts', class_labels=list('abcd'), output_scores=scores.tolist(),
                           output_classes=scores.argmax(1).tolist(), bin_id=pid.pid, roi_numbers=list(range(1, 10))))
    assert open(p).read() == want
