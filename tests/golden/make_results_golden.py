"""Generates tests/golden/dataset_golden.json and results_golden.json by running the REFERENCE'S OWN host code in the
build container: ``NeustonDataset`` (class thresholds, seeded split, class-config CSV, dataset-combining CSV ``SRC``;
neuston_data.py:20-256), ``save_run_results`` (neuston_callbacks.py:160-272) and ``SaveValidationResults``
(neuston_callbacks.py:20-156), imported unmodified from /root/reference (pyifcb / pytorch_lightning replaced by the stubs
in oracle/, h5py by a RECORDING stand-in that notes every create_dataset / attrs call, since no HDF5 library exists here).
Run once:  python tests/golden/make_results_golden.py
tests/test_results_golden.py replays the same inputs (tests/golden/results_inputs.py) through this repo's host code and
compares; /root/reference is not needed to replay.
"""
import json
import os
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ifcb_stub, ptl_stub  # noqa: E402
from tests.golden import results_inputs as ri  # noqa: E402

ifcb_stub.install()
ptl_stub.install()

# ---- recording h5py stand-in ---------------------------------------------------------------------------------------
RECORD = {}


class _Empty(object):
    def __init__(self, dtype):
        self.dtype = dtype


class _DS(object):
    def __init__(self, rec):
        self.attrs = rec['attrs']


class _File(object):
    def __init__(self, path, mode):
        self.rec = RECORD.setdefault(path, {})

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def create_dataset(self, name, data=None, compression=None, dtype=None):
        if isinstance(data, _Empty):
            rec = dict(empty=str(np.dtype(data.dtype)), attrs={})
        elif dtype == 'vlen_str':
            arr = np.asarray(data)
            rec = dict(dtype='vlen_str', compression=compression, shape=list(arr.shape),
                       data=[x.decode('utf-8') if isinstance(x, bytes) else str(x) for x in arr.reshape(-1)], attrs={})
        else:
            arr = np.asarray(data).astype(dtype)
            rec = dict(dtype=str(arr.dtype), compression=compression, shape=list(arr.shape), data=arr.astype(np.float64).reshape(-1).tolist(),
                       attrs={})
        self.rec[name] = rec
        return _DS(rec)


h5 = types.ModuleType('h5py')
h5.File, h5.Empty, h5.string_dtype = _File, _Empty, (lambda: 'vlen_str')
sys.modules['h5py'] = h5
if not hasattr(np, 'string_'):
    np.string_ = np.bytes_                      # numpy 2 removed the alias the reference uses (numpy 1.x upstream)

sys.path.insert(0, '/root/reference')
import neuston_callbacks  # noqa: E402  (the reference, unmodified)
import neuston_data  # noqa: E402


def rel(paths, root):
    return [os.path.relpath(p, root) for p in paths]


def ds_record(ds, root):
    return dict(classes=list(ds.classes), images=rel(ds.images, root), targets=[int(t) for t in ds.targets],
                ignored=[[c, n] for c, n in ds.classes_ignored_from_too_few_samples],
                limited=ds.classes_limited_from_too_many_samples, count_perclass=ds.count_perclass)


def main():
    tmp = tempfile.mkdtemp()
    ri.make_tree(tmp)
    out = {}
    # plain folder, thresholds, seeded split
    random.seed(11)
    nd = neuston_data.NeustonDataset(src=os.path.join(tmp, 'dsA'), minimum_images_per_class=2, maximum_images_per_class=6)
    out['plain'] = ds_record(nd, tmp)
    a, b = nd.split(80, 20, seed=5)
    out['split_train'], out['split_val'] = ds_record(a, tmp), ds_record(b, tmp)
    # class-config CSV
    random.seed(12)
    nd = neuston_data.NeustonDataset.from_csv(os.path.join(tmp, 'dsA'), os.path.join(tmp, 'classes.csv'), 'v2', minimum_images_per_class=2)
    out['from_csv'] = ds_record(nd, tmp)
    # dataset-combining CSV as SRC (runs from inside tmp: the CSV names the datasets by relative path)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        for name in ('combine.csv', 'combine_prio.csv'):
            random.seed(13)
            nd = neuston_data.NeustonDataset(src=name, minimum_images_per_class=1)
            out[name] = ds_record(nd, '.')
    finally:
        os.chdir(cwd)
    with open(os.path.join(HERE, 'dataset_golden.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)

    # ---- result files ----
    from scipy.io import loadmat
    res = {}
    case = ri.run_case()
    pid = ifcb_stub.Pid(case['bin'])
    pid.namespace = 'sub/'
    odir = tempfile.mkdtemp()
    for ext in ('json', 'mat', 'h5'):
        neuston_callbacks.save_run_results(case['pids'], case['scores'], case['labels'], case['timestamp'], odir,
                                           'D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.' + ext, case['model_id'], pid)
    base = os.path.join(odir, 'D2026', 'D20260102', case['bin'] + '_class')
    res['run_json'] = json.load(open(base + '.json'))
    m = loadmat(base + '.mat')
    res['run_mat'] = {k: (np.asarray(v).astype(np.float64).tolist() if np.asarray(v).dtype.kind in 'fiu' else
                          [str(x[0]) if hasattr(x, '__len__') and not isinstance(x, str) else str(x) for x in np.asarray(v).ravel()])
                      for k, v in m.items() if not k.startswith('__')}
    res['run_mat_dtypes'] = {k: str(np.asarray(v).dtype) for k, v in m.items() if not k.startswith('__')}
    res['run_h5'] = RECORD[base + '.h5']
    res['run_relpath'] = os.path.relpath(base + '.h5', odir)
    # image input, grouped by sub-directory
    icase = ri.img_case()
    neuston_callbacks.save_run_results(icase['paths'], icase['scores'], case['labels'], case['timestamp'], odir,
                                       'imgs/{INPUT_SUBDIRS}/img_results.json', case['model_id'], icase['src'])
    groups = {}
    for parent, _, files in os.walk(os.path.join(odir, 'imgs')):
        for fn in files:
            groups[os.path.relpath(os.path.join(parent, fn), odir)] = json.load(open(os.path.join(parent, fn)))
    res['img_groups'] = groups

    # ---- validation results ----
    vc = ri.val_case()

    class DS(object):
        def __init__(self, images, targets, n):
            self.images, self.targets = images, targets
            self.count_perclass = [targets.count(i) for i in range(n)]

    n = len(vc['labels'])
    train, val = DS(vc['train_images'], vc['train_targets'], n), DS(vc['val_images'], vc['val_targets'], n)
    loader = lambda ds: types.SimpleNamespace(dataset=ds)
    module = types.SimpleNamespace(current_epoch=4, hparams=types.SimpleNamespace(classes=vc['labels'], model_id='mV', cmd_timestamp='tsV'),
                                   val_dataloader=lambda: loader(val), train_dataloader=lambda: loader(train))
    vres = {}
    for ext in ('json', 'mat', 'h5'):
        trainer = types.SimpleNamespace(callback_metrics=dict(best=True, outputs=vc['scores'].copy(), input_classes=np.asarray(vc['val_targets'], dtype=np.int64),
                                                               input_srcs=list(vc['val_images'])))
        cb = neuston_callbacks.SaveValidationResults(odir, 'val/e{epoch}.' + ext, vc['series'])
        cb.on_validation_end(trainer, module)
    vres['json'] = json.load(open(os.path.join(odir, 'val', 'e4.json')))
    m = loadmat(os.path.join(odir, 'val', 'e4.mat'))
    vres['mat'] = {k: (np.asarray(v).astype(np.float64).tolist() if np.asarray(v).dtype.kind in 'fiu' else
                       [str(x[0]) if hasattr(x, '__len__') and not isinstance(x, str) else str(x) for x in np.asarray(v).ravel()])
                   for k, v in m.items() if not k.startswith('__')}
    vres['mat_dtypes'] = {k: str(np.asarray(v).dtype) for k, v in m.items() if not k.startswith('__')}
    vres['h5'] = RECORD[os.path.join(odir, 'val', 'e4.h5')]
    res['validation'] = vres
    with open(os.path.join(HERE, 'results_golden.json'), 'w') as f:
        json.dump(res, f, sort_keys=True)
    print('wrote dataset_golden.json, results_golden.json')


if __name__ == '__main__':
    main()
