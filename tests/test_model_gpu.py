"""Whole-model parity: packed ROI bytes -> preprocess kernel -> plan -> scores, against the
oracle preprocessing + torchvision fp32 (stock torch on the same GPU, TF32 off).
Gates (BASELINE.json north_star): top-1 agreement >= 99.5 % of ROIs, |score diff| <= 1e-2.
The weight fixture is stated with every number."""
import numpy as np
import pytest
import torch

from oracle.pil_resize import ref_preprocess
from tests import fixtures

pytestmark = pytest.mark.gpu


def _pipeline_scores(cuda, arch, model, imgs, img_norm, batch_cap=64, fuse=True):
    from ifcb_classifier_b200 import preprocess as pp
    from ifcb_classifier_b200.graph import CompiledNet
    net = CompiledNet(arch, model.state_dict(), batch_cap, in_kind='u8', img_norm=img_norm, device=cuda, fuse=fuse)
    scores = []
    for i in range(0, len(imgs), batch_cap):
        chunk = imgs[i:i + batch_cap]
        offs = np.cumsum([0] + [im.size for im in chunk[:-1]]).astype(np.int64)
        packed = torch.from_numpy(np.concatenate([im.reshape(-1) for im in chunk])).to(cuda)
        hs = torch.tensor([im.shape[0] for im in chunk], dtype=torch.int32, device=cuda)
        ws = torch.tensor([im.shape[1] for im in chunk], dtype=torch.int32, device=cuda)
        pp.preprocess_rois(packed, torch.from_numpy(offs).to(cuda), hs, ws, net.R, out_mode=pp.OUT_U8_GRAY,
                           out=net.inp[:len(chunk)])
        s, _, top1, top1s = net.forward(len(chunk))
        torch.cuda.synchronize()
        assert torch.equal(top1.long(), s.argmax(1))
        scores.append(s.cpu().clone())
    return torch.cat(scores)


def _ref_scores(cuda, model, x):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model.to(cuda).eval()
    out = []
    with torch.no_grad():
        for i in range(0, x.shape[0], 64):
            o = model(x[i:i + 64].to(cuda))
            out.append(torch.softmax(o, 1).cpu())
    return torch.cat(out)


@pytest.mark.parametrize('arch,n_classes', [('resnet18', 20), ('inception_v3', 20), ('resnet50', 20)])
def test_whole_model_parity_trained_fixture(cuda, arch, n_classes):
    R = 299 if arch == 'inception_v3' else 224
    norm = (([0.667] * 3), ([0.161] * 3))
    imgs, labels = fixtures.class_rois(640, n_classes, seed=1)
    x = torch.from_numpy(np.stack([ref_preprocess(im, R, norm) for im in imgs]))
    model = fixtures.ref_model(arch, n_classes)
    fixtures.calibrate_bn(model, x[:128], cuda)
    # fixture B: calibrated random init
    ref_b = _ref_scores(cuda, model, x[:256])
    got_b = _pipeline_scores(cuda, arch, model, imgs[:256], norm)
    agree_b = float((ref_b.argmax(1) == got_b.argmax(1)).float().mean())
    dmax_b = float((ref_b - got_b).abs().max())
    # fixture C: briefly trained
    fixtures.brief_train(model, x[:384], labels[:384], cuda, steps=80 if arch != 'inception_v3' else 60)
    ref_c = _ref_scores(cuda, model, x[384:])
    got_c = _pipeline_scores(cuda, arch, model, imgs[384:], norm)
    agree_c = float((ref_c.argmax(1) == got_c.argmax(1)).float().mean())
    dmax_c = float((ref_c - got_c).abs().max())
    acc = float((ref_c.argmax(1) == labels[384:]).float().mean())
    print('\n[%s] fixture B (calibrated random init): top-1 agreement %.4f, max|dscore| %.3e ; '
          'fixture C (briefly trained, ref acc %.2f, mean max-prob %.2f): top-1 agreement %.4f, max|dscore| %.3e'
          % (arch, agree_b, dmax_b, acc, float(ref_c.max(1).values.mean()), agree_c, dmax_c))
    assert dmax_b <= 1e-2 and dmax_c <= 1e-2
    assert agree_c >= 0.995
