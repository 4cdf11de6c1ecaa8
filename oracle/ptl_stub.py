"""Tiny stand-in for the handful of pytorch_lightning names the reference's
``neuston_models.py`` needs at import time, so that ``NeustonModel`` can be
imported UNMODIFIED from /root/reference for golden-vector generation
(tests/golden/make_golden.py; build container only).  TEST INFRASTRUCTURE.
pytorch-lightning 1.3.8 is pinned upstream and absent from this image."""
import sys
import types

import torch.nn as nn


class LightningModule(nn.Module):
    def save_hyperparameters(self, hparams):
        self.hparams = hparams

    def log(self, *a, **k):
        pass


def install():
    ptl = types.ModuleType('pytorch_lightning')
    ptl.LightningModule = LightningModule
    cb = types.ModuleType('pytorch_lightning.callbacks')
    base = types.ModuleType('pytorch_lightning.callbacks.base')

    class Callback(object):
        pass
    base.Callback = Callback
    cb.base = base
    ptl.callbacks = cb
    sys.modules.update({'pytorch_lightning': ptl, 'pytorch_lightning.callbacks': cb,
                        'pytorch_lightning.callbacks.base': base})
    return ptl
