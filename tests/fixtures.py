"""Weight / data fixtures for the whole-model parity tests.

A = raw random init (degenerate for Inception-v3: BN is identity, logits explode)
B = random init + BN running statistics calibrated on synthetic ROIs (near-uniform
    softmax, adversarial for low precision)
C = B + a short Adam run on synthetic separable classes (representative of a
    trained checkpoint: confident predictions)
The reference model is built exactly as get_namebrand_model does
(reference neuston_models.py:22-45) from torchvision.
"""
import warnings

import numpy as np
import torch
import torch.nn as nn


def ref_model(name, n_classes, seed=0):
    """get_namebrand_model(name, n_classes, pretrained=False) restated."""
    import torchvision.models as M
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if name == 'inception_v3':
            m = M.inception_v3(weights=None, aux_logits=True, init_weights=True)
            m.AuxLogits.fc = nn.Linear(m.AuxLogits.fc.in_features, n_classes)
            m.fc = nn.Linear(m.fc.in_features, n_classes)
        elif name.startswith('resnet'):
            m = getattr(M, name)(weights=None)
            m.fc = nn.Linear(m.fc.in_features, n_classes)
        elif name == 'squeezenet':
            m = M.squeezenet1_1(weights=None)
            m.classifier[1] = nn.Conv2d(512, n_classes, kernel_size=(1, 1), stride=(1, 1))
            m.num_classes = n_classes
        elif name.startswith('densenet'):
            m = getattr(M, name)(weights=None)
            m.classifier = nn.Linear(m.classifier.in_features, n_classes)
        elif name == 'alexnet' or name.startswith('vgg'):
            m = getattr(M, name)(weights=None)
            m.classifier[6] = nn.Linear(m.classifier[6].in_features, n_classes)
        else:
            raise KeyError('model unknown!')
    return m


def class_rois(n, n_classes, seed=0, hw=None):
    """Synthetic separable classes: oriented gratings whose frequency encodes the label.
    Returns (list of uint8[h,w], labels int64[n])."""
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, n_classes, n)
    imgs = []
    for i in range(n):
        if hw is None:
            h = int(np.clip(rng.lognormal(np.log(60.0), 0.5), 16, 400))
            w = int(np.clip(rng.lognormal(np.log(90.0), 0.5), 16, 500))
        else:
            h, w = hw
        k = int(labels[i])
        fx, fy = 0.5 + (k % 5) * 1.5, 0.5 + (k // 5) * 1.5
        yy, xx = np.mgrid[0:h, 0:w]
        ph = rng.uniform(0, 2 * np.pi)
        img = 150 + 70 * np.sin(2 * np.pi * (fx * xx / w + fy * yy / h) + ph) + rng.normal(0, 8, (h, w))
        imgs.append(np.clip(img, 0, 255).astype(np.uint8))
    return imgs, torch.from_numpy(labels)


def calibrate_bn(model, x, device):
    """Fixture B: set BN running stats to the batch statistics of ``x`` (momentum=None)."""
    model.to(device).train()
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.reset_running_stats()
            m.momentum = None
    with torch.no_grad():
        for i in range(0, x.shape[0], 32):
            model(x[i:i + 32].to(device))
    model.eval()
    return model


def brief_train(model, x, y, device, steps=60, batch=32, lr=1e-3):
    """Fixture C: a few Adam steps, loss as NeustonModel.loss (neuston_models.py:70-78)."""
    model.to(device).train()
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.momentum = 0.1
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    crit = nn.CrossEntropyLoss()
    n = x.shape[0]
    g = torch.Generator().manual_seed(0)
    for s in range(steps):
        idx = torch.randint(0, n, (batch,), generator=g)
        xb, yb = x[idx].to(device), y[idx].to(device)
        out = model(xb)
        if isinstance(out, tuple) and len(out) == 2:
            loss = crit(out[0], yb) + 0.4 * crit(out[1], yb)
        else:
            loss = crit(out, yb)
        opt.zero_grad()
        loss.backward()
        opt.step()
    model.eval()
    return model
