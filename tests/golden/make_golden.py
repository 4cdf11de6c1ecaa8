"""Generates tests/golden/*.npz by running the REFERENCE'S OWN code in the build
container: ``IfcbBinDataset`` and ``NeustonModel`` are imported unmodified from
/root/reference (pyifcb / pytorch_lightning replaced by the stubs in oracle/).
Run once:  python tests/golden/make_golden.py
The vectors pin the oracle (tests/test_oracle_golden.py); /root/reference is not
needed to replay them.
"""
import argparse
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ifcb_stub, ptl_stub, synth_bins  # noqa: E402

ifcb_stub.install()
ptl_stub.install()
sys.path.insert(0, '/root/reference')
import neuston_data  # noqa: E402  (the reference, unmodified)
import neuston_models  # noqa: E402

DIMS = [(60, 90), (16, 16), (33, 200), (299, 299), (224, 224), (300, 120), (500, 3), (1034, 8),
        (225, 2), (17, 1380), (1034, 1380), (120, 97), (1, 40), (41, 1)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    # ---- preprocessing golden: the reference dataset on a synthetic bin ----
    b = synth_bins.make_bin(7, dims=DIMS, variant='uniform', empty_every=5)
    tmp = tempfile.mkdtemp()
    synth_bins.write_bin(tmp, b)
    fb = list(ifcb_stub.DataDirectory(tmp))[0]
    out = dict(adc=b['adc'], roi=b['roi'], lid=np.array(b['lid']))
    for R in (299, 224):
        for tag, norm in (('plain', None), ('norm', ['0.667', '0.161']), ('norm3', ['0.5,0.4,0.3', '0.2,0.25,0.3'])):
            ds = neuston_data.IfcbBinDataset(fb, R, norm)
            hashes, grays = [], []
            for i in range(len(ds)):
                t, pid = ds[i]
                hashes.append(sha(t.numpy()))
                if tag == 'plain':
                    g = np.rint(t[0].numpy() * 255).astype(np.uint8)
                    assert np.array_equal(g.astype(np.float32) / np.float32(255), t[0].numpy())
                    grays.append(g)
            out['sha_%d_%s' % (R, tag)] = np.array(hashes)
            if grays:
                out['gray_%d' % R] = np.stack(grays)
            out['pids'] = np.array(ds.pids)
    np.savez_compressed(os.path.join(HERE, 'preprocess_golden.npz'), **out)

    # ---- model golden: the reference NeustonModel.test_step on fixed seeds ----
    mout = {}
    for name, R, C in (('resnet18', 224, 10), ('inception_v3', 299, 10)):
        torch.manual_seed(0)
        hp = argparse.Namespace(MODEL=name, classes=['c%d' % i for i in range(C)], pretrained=False)
        m = neuston_models.NeustonModel(hp)
        m.eval()
        g = torch.Generator().manual_seed(1)
        x = torch.rand(3, 3, R, R, generator=g)
        with torch.no_grad():
            o = m.test_step((x, ['a', 'b', 'c']), 0)['test_outputs']
        mout[name + '_scores'] = o.numpy()
        mout[name + '_keys'] = np.array(sorted(m.model.state_dict().keys()))
        tgt = torch.tensor([1, 2, 3])
        m.train()
        torch.manual_seed(2)
        lo = m.loss(tgt, m.forward(x))
        mout[name + '_train_loss'] = np.array(float(lo))
    np.savez_compressed(os.path.join(HERE, 'model_golden.npz'), **mout)
    print('wrote golden vectors')


if __name__ == '__main__':
    main()
