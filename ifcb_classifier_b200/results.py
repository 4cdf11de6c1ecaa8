"""Per-bin result files with the reference's layout (neuston_callbacks.py:160-272).

``save_run_results`` mirrors argument names and the templated path
(``{BIN_ID} {BIN_YEAR} {BIN_DATE} {INPUT_SUBDIRS}``).  Formats: ``.h5`` (the reference's default; written by
the in-tree ``h5lite`` module, no h5py needed), ``.mat`` and ``.json``.  The top-1 index / winning score come from the GPU
head kernel; they equal ``argmax`` / ``max`` of the score rows by construction.
"""
import json
import os

import numpy as np


def outfile_path(outdir, outfile, pid):
    d = dict(BIN_ID=pid.pid, INPUT_SUBDIRS=pid.namespace or '', BIN_YEAR=pid.year, BIN_DATE=pid.yearday)
    return os.path.join(outdir, outfile).format(**d).replace(2 * os.sep, os.sep)


def save_run_results(input_images, output_scores, class_labels, timestamp, outdir, outfile, model_id=None,
                     input_obj=None, output_classes=None):
    output_scores = np.asarray(output_scores)
    if output_classes is None:
        output_classes = np.argmax(output_scores, axis=1)
    assert output_scores.shape[0] == len(output_classes), 'wrong number inputs-to-outputs'
    assert output_scores.shape[1] == len(class_labels), 'wrong number of class labels'
    results = dict(version='v3', model_id=model_id, timestamp=timestamp, class_labels=list(class_labels),
                   input_images=list(input_images), output_classes=np.asarray(output_classes),
                   output_scores=output_scores)
    if hasattr(input_obj, 'pid'):                                  # a bin (ifcb.Pid upstream)
        results['bin_id'] = input_obj.pid
        results['roi_numbers'] = [int(str(img).rsplit('_', 1)[1]) for img in input_images]
        path = outfile_path(outdir, outfile, input_obj)
        os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
        _save(path, results)
        return path
    # --type img (neuston_callbacks.py:186-206): one file, or one per sub-directory of the input tree
    outfile = os.path.join(outdir, outfile)
    if '{INPUT_SUBDIRS}' in outfile:
        input_src = input_obj if (isinstance(input_obj, str) and os.path.isdir(input_obj)) else ''
        groups = {}
        for path_, cls_, sc_ in zip(input_images, results['output_classes'], output_scores):
            parent = os.path.dirname(path_.replace(input_src, ''))
            g = groups.setdefault(parent, dict(results, input_images=[], output_classes=[], output_scores=[]))
            g['input_images'].append(os.path.basename(path_))
            g['output_classes'].append(cls_)
            g['output_scores'].append(sc_)
        written = []
        for parent, g in groups.items():
            sub = outfile.format(INPUT_SUBDIRS=parent)
            os.makedirs(os.path.dirname(sub) or '.', exist_ok=True)
            g['output_classes'] = np.asarray(g['output_classes'], dtype=results['output_classes'].dtype)
            g['output_scores'] = np.asarray(g['output_scores'], dtype=output_scores.dtype)
            _save(sub, g)
            written.append(sub)
        return written
    os.makedirs(os.path.dirname(outfile) or '.', exist_ok=True)
    _save(outfile, results)
    return outfile


_SCORES_MARK = '@@ifcb_output_scores@@'


def _scores_json(scores):
    from . import _lib
    a = np.ascontiguousarray(scores, dtype=np.float32)
    if a.ndim != 2:
        return json.dumps(a.tolist())
    cap = a.shape[0] * (a.shape[1] * 28 + 4) + 4
    buf = np.empty(cap, np.uint8)
    n = _lib.lib().ifcb_format_scores_json(a.ctypes.data, a.shape[0], a.shape[1], buf.ctypes.data, cap)
    if n < 0:
        _lib.check(-1, 'format_scores_json')
    return buf[:n].tobytes().decode('ascii')


def _save(path, r):
    ext = os.path.splitext(path)[-1]
    assert ext in ['.json', '.mat', '.h5'], 'output fileformat "{}" not valid'.format(ext)
    if ext == '.json':
        out = dict(version=r['version'], model_id=r['model_id'], timestamp=r['timestamp'],
                   class_labels=r['class_labels'], output_scores=_SCORES_MARK,
                   output_classes=[int(c) for c in r['output_classes']])
        if 'bin_id' in r:
            out['bin_id'], out['roi_numbers'] = r['bin_id'], r['roi_numbers']
        else:
            out['input_images'] = r['input_images']
        # the score matrix (2048 x 100 numbers per bin) is formatted by the library, byte-identical to json.dumps
        # (ifcb_format_scores_json: ~30x faster than Python's encoder and outside the GIL); the rest is json.dumps
        text = json.dumps(out).replace('"%s"' % _SCORES_MARK, _scores_json(r['output_scores']), 1)
        with open(path, 'w') as f:
            f.write(text)
    elif ext == '.mat':
        from scipy.io import savemat
        out = dict(output_classes=r['output_classes'].astype('u4') + 1,      # matlab is 1-based
                   version=r['version'], model_id=r['model_id'] if r['model_id'] is not None else '',
                   timestamp=r['timestamp'], output_scores=r['output_scores'].astype('f4'),
                   class_labels=np.asarray(r['class_labels'], dtype='object'))
        if 'bin_id' in r:
            out['bin_id'], out['roi_numbers'] = r['bin_id'], r['roi_numbers']
        else:
            out['input_images'] = np.asarray(r['input_images'], dtype='object')
        savemat(path, out, do_compression=True)
    else:
        # _save_run_results_hdf (neuston_callbacks.py:252-268): same datasets, dtypes, gzip filter and attributes, written
        # by the in-tree HDF5 writer (h5lite: h5py / libhdf5 are not needed)
        from . import h5lite
        attrs = dict(version=r['version'], model_id=r['model_id'] if r['model_id'] is not None else '', timestamp=r['timestamp'])
        ds = dict(output_classes=dict(data=r['output_classes'], dtype='float16'),
                  output_scores=dict(data=r['output_scores'], dtype='float16'),
                  class_labels=dict(data=[str(c) for c in r['class_labels']], dtype='vlen_str'))
        if 'bin_id' in r:
            attrs['bin_id'] = str(r['bin_id'])
            ds['roi_numbers'] = dict(data=r['roi_numbers'], dtype='uint16')
        else:
            ds['input_images'] = dict(data=[str(p_) for p_ in r['input_images']], dtype='vlen_str')
        ds['metadata'] = dict(data=h5lite.Empty('f'), attrs=attrs)
        h5lite.write(path, ds)
