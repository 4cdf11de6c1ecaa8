"""Per-bin result files with the reference's layout (neuston_callbacks.py:160-272).

``save_run_results`` mirrors argument names and the templated path
(``{BIN_ID} {BIN_YEAR} {BIN_DATE} {INPUT_SUBDIRS}``).  Formats: ``.json`` and ``.mat``
always; ``.h5`` needs ``h5py`` (absent from this image -- a clear error is raised instead
of silently writing something else).  The top-1 index / winning score come from the GPU
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
    results['bin_id'] = input_obj.pid
    results['roi_numbers'] = [int(str(img).rsplit('_', 1)[1]) for img in input_images]
    path = outfile_path(outdir, outfile, input_obj)
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    _save(path, results)
    return path


def _save(path, r):
    ext = os.path.splitext(path)[-1]
    assert ext in ['.json', '.mat', '.h5'], 'output fileformat "{}" not valid'.format(ext)
    if ext == '.json':
        out = dict(version=r['version'], model_id=r['model_id'], timestamp=r['timestamp'],
                   class_labels=r['class_labels'], output_scores=r['output_scores'].tolist(),
                   output_classes=[int(c) for c in r['output_classes']], bin_id=r['bin_id'],
                   roi_numbers=r['roi_numbers'])
        with open(path, 'w') as f:
            json.dump(out, f)
    elif ext == '.mat':
        from scipy.io import savemat
        out = dict(output_classes=r['output_classes'].astype('u4') + 1,      # matlab is 1-based
                   version=r['version'], model_id=r['model_id'] if r['model_id'] is not None else '',
                   timestamp=r['timestamp'], output_scores=r['output_scores'].astype('f4'),
                   class_labels=np.asarray(r['class_labels'], dtype='object'), bin_id=r['bin_id'],
                   roi_numbers=r['roi_numbers'])
        savemat(path, out, do_compression=True)
    else:
        try:
            import h5py as h5
        except ImportError:
            raise RuntimeError('writing %s needs h5py, which is not installed; use --outfile with .mat or .json' % path)
        with h5.File(path, 'w') as f:
            meta = f.create_dataset('metadata', data=h5.Empty('f'))
            meta.attrs['version'], meta.attrs['model_id'] = r['version'], r['model_id']
            meta.attrs['timestamp'], meta.attrs['bin_id'] = r['timestamp'], r['bin_id']
            f.create_dataset('output_classes', data=r['output_classes'], compression='gzip', dtype='float16')
            f.create_dataset('output_scores', data=r['output_scores'], compression='gzip', dtype='float16')
            f.create_dataset('class_labels', data=np.bytes_(r['class_labels']), compression='gzip',
                             dtype=h5.string_dtype())
            f.create_dataset('roi_numbers', data=r['roi_numbers'], compression='gzip', dtype='uint16')
