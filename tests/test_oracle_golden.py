"""CPU: the oracle replays the golden vectors recorded from the reference's own code
(tests/golden/make_golden.py ran /root/reference's IfcbBinDataset / NeustonModel unmodified)."""
import hashlib
import os

import numpy as np
import torch

from oracle import ifcb_stub, model_ref, synth_bins
from oracle.pil_resize import ref_preprocess, resize_gray_u8, bilinear_coeffs

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden_bin(tmp_path):
    g = np.load(os.path.join(GOLD, 'preprocess_golden.npz'))
    b = dict(lid=str(g['lid']), adc=g['adc'], roi=g['roi'])
    synth_bins.write_bin(str(tmp_path), b)
    fb = list(ifcb_stub.DataDirectory(str(tmp_path)))[0]
    return g, fb


def test_preprocess_oracle_matches_reference_dataset(tmp_path):
    g, fb = _golden_bin(tmp_path)
    imgs = list(fb.images.items())
    assert [fb.pid.with_target(t) for t, _ in imgs] == list(g['pids'])
    for R in (299, 224):
        for tag, norm in (('plain', None), ('norm', ['0.667', '0.161']), ('norm3', ['0.5,0.4,0.3', '0.2,0.25,0.3'])):
            want = g['sha_%d_%s' % (R, tag)]
            for i, (t, im) in enumerate(imgs):
                assert _sha(ref_preprocess(im, R, norm)) == want[i], (R, tag, im.shape)
        for i, (t, im) in enumerate(imgs):
            assert np.array_equal(resize_gray_u8(im, R), g['gray_%d' % R][i])


def test_coeffs_are_normalised_fixed_point():
    for (n_in, n_out) in [(90, 299), (1380, 299), (299, 299), (16, 224), (1034, 224), (1, 299)]:
        xmin, cnt, kk = bilinear_coeffs(n_in, n_out)
        assert kk.shape[1] == int(np.ceil(max(n_in / n_out, 1.0))) * 2 + 1
        assert np.all(np.abs(kk.sum(1) - (1 << 22)) <= kk.shape[1])      # rows sum to ~1.0 in Q22
        assert np.all(xmin >= 0) and np.all(xmin + cnt <= n_in) and np.all(cnt >= 1)


def test_skipped_trigger_rows_and_pids(tmp_path):
    b = synth_bins.make_bin(11, n_rois=20, empty_every=4)
    synth_bins.write_bin(str(tmp_path), b)
    fb = list(ifcb_stub.DataDirectory(str(tmp_path)))[0]
    assert list(fb.images.keys()) == sorted(b['images'].keys())          # zero-area rows skipped, 1-based targets kept
    for t, im in fb.images.items():
        assert np.array_equal(im, b['images'][t])
    p = ifcb_stub.Pid(fb.pid.with_target(17))
    assert p.target == 17 and p.bin_lid == b['lid'] and p.year == '2026'


def test_model_oracle_matches_reference_neustonmodel():
    g = np.load(os.path.join(GOLD, 'model_golden.npz'))
    for name, R, C in (('resnet18', 224, 10), ('inception_v3', 299, 10)):
        torch.manual_seed(0)
        m = model_ref.get_namebrand_model(name, C, False)
        assert sorted(m.state_dict().keys()) == list(g[name + '_keys'])
        x = torch.rand(3, 3, R, R, generator=torch.Generator().manual_seed(1))
        s = model_ref.test_step_scores(m, x).numpy()
        assert np.allclose(s, g[name + '_scores'], rtol=1e-4, atol=1e-6), name
        m.train()
        torch.manual_seed(2)
        lo = float(model_ref.loss(m(x), torch.tensor([1, 2, 3])))
        assert abs(lo - float(g[name + '_train_loss'])) < 1e-3 * abs(lo), name
    try:
        model_ref.get_namebrand_model('nope', 3)
        assert False
    except KeyError as e:
        assert 'model unknown!' in str(e)
