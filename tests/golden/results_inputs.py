"""Deterministic inputs shared by tests/golden/make_results_golden.py (which feeds them to the reference's own host code)
and tests/test_results_golden.py (which feeds them to this repo's host code)."""
import os

import numpy as np


def make_tree(root):
    """Two class-per-folder datasets of empty image files, a class-config CSV and two dataset-combining CSVs."""
    spec = {
        'dsA': dict(Akashiwo=9, Bacillaria=4, Bidulphia=3, Cochlodinium=7, Didinium_sp=1, Ephemera=5),
        'dsB': dict(Akashiwo=3, Bacillaria=2, Ceratium=6, Cochlodinium=2),
        'dsC': dict(Akashiwo=2, Ceratium=3),
    }
    for ds, classes in spec.items():
        for c, n in classes.items():
            os.makedirs(os.path.join(root, ds, c), exist_ok=True)
            for i in range(n):
                open(os.path.join(root, ds, c, '%s_%s_%03d.png' % (ds, c[:3], (i * 7) % 11 * 10 + i)), 'w').close()
            open(os.path.join(root, ds, c, 'notes.txt'), 'w').close()             # not an image: ignored
            open(os.path.join(root, ds, c, 'UPPER_%s.PNG' % c[:2]), 'w').close()  # upper-case extension: ignored upstream
    with open(os.path.join(root, 'classes.csv'), 'w') as f:
        f.write('class,v1,v2\nAkashiwo,1,1\nBacillaria,1,0\nBidulphia,1,BIDOUF\nCochlodinium,1,BIDOUF\nDidinium_sp,1,1\nEphemera,0,1\nMissing,1,1\n')
    with open(os.path.join(root, 'combine.csv'), 'w') as f:
        f.write('class,dsA,dsB\nAkashiwo,1,1\nBacillaria,0,1\nBidulphia,BIDOUF,1\nCochlodinium,BIDOUF,0\nCeratium,1,1\nEphemera,1,1\n')
    with open(os.path.join(root, 'combine_prio.csv'), 'w') as f:
        f.write('class,2:dsA,1:dsB,dsC\nAkashiwo,1,1,1\nBacillaria,1,1,1\nCeratium,1,1,1\nCochlodinium,1,0,1\n')


def run_case():
    rng = np.random.default_rng(42)
    scores = rng.random((23, 5)).astype(np.float32)
    scores /= scores.sum(1, keepdims=True)
    lid = 'D20260102T030405_IFCB999'
    targets = [1, 2, 3, 5, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21, 22, 23, 24, 25, 26, 300]
    return dict(bin=lid, pids=['%s_%05d' % (lid, t) for t in targets], scores=scores, labels=['a', 'b', 'c_long_name', 'd', 'e'],
                timestamp='2026-01-02T03:04:05+00:00', model_id='golden_model')


def img_case():
    rng = np.random.default_rng(43)
    scores = rng.random((6, 5)).astype(np.float32)
    scores /= scores.sum(1, keepdims=True)
    src = '/data/imgs/'
    paths = [src + 'x/a.png', src + 'x/b.png', src + 'y/z/c.png', src + 'd.png', src + 'y/z/e.png', src + 'x/f.png']
    return dict(src=src, paths=paths, scores=scores)


def val_case():
    rng = np.random.default_rng(44)
    labels = ['k0', 'k1', 'k2', 'k3']
    val_targets = [0, 0, 1, 1, 1, 2, 2, 3, 3, 3, 3, 0]
    scores = rng.random((len(val_targets), 4)).astype(np.float32)
    for i, t in enumerate(val_targets):
        if i % 4 != 3:
            scores[i, t] += 1.0                          # mostly right
    scores /= scores.sum(1, keepdims=True)
    train_targets = [0] * 5 + [1] * 4 + [2] * 6 + [3] * 3
    series = ('image_fullpaths image_basenames training_image_fullpaths training_image_basenames training_classes output_winscores '
              'output_scores confusion_matrix counts_perclass val_counts_perclass train_counts_perclass f1_perclass f1_weighted f1_macro '
              'recall_perclass recall_macro precision_weighted classes_by_f1 classes_by_recall classes_by_count').split()
    return dict(labels=labels, val_targets=val_targets, scores=scores,
                val_images=['/d/v/%s/v%02d.png' % (labels[t], i) for i, t in enumerate(val_targets)],
                train_targets=train_targets, train_images=['/d/t/%s/t%02d.png' % (labels[t], i) for i, t in enumerate(train_targets)],
                series=series)
