"""CPU: this repo's host code against vectors recorded from the REFERENCE'S OWN host code run in the build container
(tests/golden/make_results_golden.py imports neuston_data / neuston_callbacks unmodified from the reference):
datasets (thresholds, seeded split, class-config CSV, dataset-combining CSV ``SRC``), per-bin result files
(.json / .mat / .h5) and the validation results files."""
import argparse
import json
import os
import random

import numpy as np
import pytest

from tests.golden import results_inputs as ri

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden(name):
    with open(os.path.join(HERE, 'golden', name)) as f:
        return json.load(f)


def _rec(ds, root):
    return dict(classes=list(ds.classes), images=[os.path.relpath(p, root) for p in ds.images], targets=[int(t) for t in ds.targets],
                ignored=[[c, n] for c, n in ds.classes_ignored_from_too_few_samples],
                limited=ds.classes_limited_from_too_many_samples, count_perclass=ds.count_perclass)


def test_datasets_match_the_reference(tmp_path, monkeypatch):
    """NeustonDataset / split / from_csv / CSV-as-SRC (reference neuston_data.py:20-256) -- same classes, same images in the
    same order, same targets, for the same `random` seeds."""
    from ifcb_classifier_b200.neuston_data import NeustonDataset
    g = _golden('dataset_golden.json')
    tmp = str(tmp_path)
    ri.make_tree(tmp)
    random.seed(11)
    nd = NeustonDataset(src=os.path.join(tmp, 'dsA'), minimum_images_per_class=2, maximum_images_per_class=6)
    assert _rec(nd, tmp) == g['plain']
    a, b = nd.split(80, 20, seed=5)
    assert _rec(a, tmp) == g['split_train'] and _rec(b, tmp) == g['split_val']
    random.seed(12)
    nd = NeustonDataset.from_csv(os.path.join(tmp, 'dsA'), os.path.join(tmp, 'classes.csv'), 'v2', minimum_images_per_class=2)
    assert _rec(nd, tmp) == g['from_csv']
    monkeypatch.chdir(tmp)
    for name in ('combine.csv', 'combine_prio.csv'):
        random.seed(13)
        nd = NeustonDataset(src=name, minimum_images_per_class=1)
        assert _rec(nd, '.') == g[name], name


def _mat_dict(path):
    from scipy.io import loadmat
    m = loadmat(path)
    vals = {k: (np.asarray(v).astype(np.float64).tolist() if np.asarray(v).dtype.kind in 'fiu' else
                [str(x[0]) if hasattr(x, '__len__') and not isinstance(x, str) else str(x) for x in np.asarray(v).ravel()])
            for k, v in m.items() if not k.startswith('__')}
    return vals, {k: str(np.asarray(v).dtype) for k, v in m.items() if not k.startswith('__')}


def _check_h5(path, want):
    """``want``: what the reference asked h5py to write (recorded call by call)."""
    from ifcb_classifier_b200 import h5lite
    got = h5lite.read(path)
    assert sorted(got) == sorted(want)
    for name, w in want.items():
        d = got[name]
        if 'empty' in w:
            assert d.shape is None and str(np.dtype(d.dtype)) == w['empty']
        elif w['dtype'] == 'vlen_str':
            assert d.dtype == 'vlen_str' and list(d.shape) == w['shape'] and d.data.reshape(-1).tolist() == w['data']
        else:
            assert str(d.dtype) == w['dtype'] and list(d.shape) == w['shape'], name
            assert d.data.astype(np.float64).reshape(-1).tolist() == w['data'], name
        if 'empty' not in w:
            assert (w['compression'] == 'gzip') == (d.filters == [(1, (4,))]), name
        assert sorted(d.attrs) == sorted(w['attrs']), name
        for k, v in w['attrs'].items():
            assert d.attrs[k] == v or (isinstance(v, float) and abs(d.attrs[k] - v) < 1e-12), (name, k)


def test_run_result_files_match_the_reference(tmp_path):
    """save_run_results (reference neuston_callbacks.py:160-272): path template, .json content, .mat variables and dtypes,
    .h5 datasets / dtypes / attributes; --type img grouping by INPUT_SUBDIRS."""
    from ifcb_classifier_b200 import ifcb_io, results
    g = _golden('results_golden.json')
    case = ri.run_case()
    pid = ifcb_io.Pid(case['bin'])
    pid.namespace = 'sub/'
    odir = str(tmp_path)
    paths = {}
    for ext in ('json', 'mat', 'h5'):
        paths[ext] = results.save_run_results(case['pids'], case['scores'], case['labels'], case['timestamp'], odir,
                                              'D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.' + ext, case['model_id'], pid)
    assert os.path.relpath(paths['h5'], odir) == g['run_relpath']
    assert json.load(open(paths['json'])) == g['run_json']
    vals, dtypes = _mat_dict(paths['mat'])
    assert vals == g['run_mat'] and dtypes == g['run_mat_dtypes']
    _check_h5(paths['h5'], g['run_h5'])
    icase = ri.img_case()
    written = results.save_run_results(icase['paths'], icase['scores'], case['labels'], case['timestamp'], odir,
                                       'imgs/{INPUT_SUBDIRS}/img_results.json', case['model_id'], icase['src'])
    got = {os.path.relpath(p, odir): json.load(open(p)) for p in written}
    assert got == g['img_groups']


def test_validation_result_files_match_the_reference(tmp_path):
    """save_validation_results vs SaveValidationResults.on_validation_end (reference neuston_callbacks.py:20-156): the chosen
    series, F1 / recall / precision, confusion matrix, class orderings; .json / .mat / .h5."""
    from ifcb_classifier_b200.train_loop import save_validation_results
    g = _golden('results_golden.json')['validation']
    vc = ri.val_case()

    class DS(object):
        def __init__(self, images, targets, n):
            self.images, self.targets = images, targets
            self.count_perclass = [targets.count(i) for i in range(n)]

    n = len(vc['labels'])
    train, val = DS(vc['train_images'], vc['train_targets'], n), DS(vc['val_images'], vc['val_targets'], n)
    args = argparse.Namespace(classes=vc['labels'], model_id='mV', cmd_timestamp='tsV', outdir=str(tmp_path))
    out = {}
    for ext in ('json', 'mat', 'h5'):
        out[ext] = save_validation_results('val/e{epoch}.' + ext, vc['series'], args, 4, train, val, np.asarray(vc['val_targets'], dtype=np.int64),
                                           vc['scores'].copy(), list(vc['val_images']))
    assert out['json'].endswith('val/e4.json')
    assert json.load(open(out['json'])) == g['json']
    vals, dtypes = _mat_dict(out['mat'])
    assert sorted(vals) == sorted(g['mat'])
    for k in vals:
        assert dtypes[k] == g['mat_dtypes'][k], k
        assert vals[k] == g['mat'][k] or np.allclose(np.asarray(vals[k], np.float64), np.asarray(g['mat'][k], np.float64), rtol=1e-6), k
    _check_h5(out['h5'], g['h5'])
