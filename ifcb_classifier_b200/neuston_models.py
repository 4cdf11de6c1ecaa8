"""Model registry and module hooks of the reference, re-hosted on the B200 plan.

Mirrors ``/root/reference/neuston_models.py``: ``get_namebrand_model`` (22-45: same
model names, same head swap, ``KeyError("model unknown!")``), ``NeustonModel`` hooks
(48-190: ``forward``, ``test_step``, ``RunResults``, hparams with ``MODEL``,
``classes``, ``pretrained``, ``resize``, ``img_norm``, ``model_id`` ...).
pytorch_lightning is not used: the loop runs in ``engine.BinClassifier`` / ``neuston_net``.
Weights live in a torchvision-layout ``state_dict`` (keys prefixed ``model.`` inside a
checkpoint, exactly like a Lightning ``.ptl``); the forward pass is the compiled sm_100a
plan -- torchvision is only used to *initialise / name* parameters, never to compute.
"""
import argparse
import warnings

import torch
import torch.nn as nn

from .graph import CompiledNet, RESNET_CFG, PLAIN_ARCHS, DENSE_ARCHS

ACCELERATED = ('inception_v3',) + tuple(RESNET_CFG) + tuple(PLAIN_ARCHS) + ('squeezenet',) + tuple(DENSE_ARCHS)


def _torchvision_module(model_name, num_o_classes, pretrained):
    import torchvision.models as M
    weights = 'DEFAULT' if pretrained else None     # needs the torchvision weight cache (no network here)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if model_name == 'inception_v3':
            m = M.inception_v3(weights=weights, aux_logits=True, **({} if pretrained else {'init_weights': True}))
            m.AuxLogits.fc = nn.Linear(m.AuxLogits.fc.in_features, num_o_classes)
            m.fc = nn.Linear(m.fc.in_features, num_o_classes)
        elif model_name == 'alexnet':
            m = M.alexnet(weights=weights)
            m.classifier[6] = nn.Linear(m.classifier[6].in_features, num_o_classes)
        elif model_name == 'squeezenet':
            m = M.squeezenet1_1(weights=weights)
            m.classifier[1] = nn.Conv2d(512, num_o_classes, kernel_size=(1, 1), stride=(1, 1))
            m.num_classes = num_o_classes
        elif model_name.startswith('vgg'):
            m = getattr(M, model_name)(weights=weights)
            m.classifier[6] = nn.Linear(m.classifier[6].in_features, num_o_classes)
        elif model_name.startswith('resnet'):
            m = getattr(M, model_name)(weights=weights)
            m.fc = nn.Linear(m.fc.in_features, num_o_classes)
        elif model_name.startswith('densenet'):
            m = getattr(M, model_name)(weights=weights)
            m.classifier = nn.Linear(m.classifier.in_features, num_o_classes)
        else:
            raise KeyError("model unknown!")
    return m


class B200Model(object):
    """A named backbone: torchvision-layout parameters + compiled B200 plans."""

    def __init__(self, model_name, num_o_classes, pretrained=False, state_dict=None):
        self.model_name, self.num_classes, self.pretrained = model_name, int(num_o_classes), bool(pretrained)
        # torchvision's factory sets transform_input=True exactly when weights are requested
        # (inception.py:463-465); it is a module attribute, not part of the state_dict.
        self.transform_input = bool(pretrained) and model_name == 'inception_v3'
        if state_dict is None:
            try:
                state_dict = _torchvision_module(model_name, num_o_classes, pretrained).state_dict()
            except KeyError:
                raise
            except Exception as e:
                if pretrained:
                    raise RuntimeError('%s: pretrained torchvision weights are not available on this host (%s: %s); '
                                       'pass --untrain to start from random init' % (model_name, type(e).__name__, e))
                raise
        self._sd = {k: v.detach().clone() for k, v in state_dict.items()}
        self._plans = {}

    def state_dict(self):
        return self._sd

    def load_state_dict(self, sd):
        missing = [k for k in self._sd if k not in sd]
        if missing:
            raise KeyError('missing keys in state_dict: %s ...' % missing[:3])
        self._sd = {k: sd[k].detach().clone() for k in self._sd}
        self._plans = {}

    def compile(self, batch_cap, in_kind='u8', img_norm=None, device='cuda', dtype='fp16'):
        if self.model_name not in ACCELERATED:
            raise NotImplementedError('%s resolves as a model name but has no B200 plan yet '
                                      '(accelerated: %s)' % (self.model_name, ', '.join(ACCELERATED)))
        key = (int(batch_cap), in_kind, repr(img_norm), str(device), dtype)
        if key not in self._plans:
            self._plans[key] = CompiledNet(self.model_name, self._sd, batch_cap, in_kind=in_kind, img_norm=img_norm,
                                           transform_input=self.transform_input, device=device, dtype=dtype)
        return self._plans[key]

    def __call__(self, x):
        """Eval-mode forward of a float32 [B,3,R,R] CUDA tensor -> logits float32 [B, C]."""
        if not x.is_cuda:
            raise RuntimeError('B200Model: input must be a CUDA tensor (there is no CPU path)')
        B = int(x.shape[0])
        cap = max(16, 1 << (B - 1).bit_length())
        net = self.compile(cap, in_kind='f32', device=x.device)
        net.inp[:B].copy_(x)
        _, logits, _, _ = net.forward(B)
        return logits.clone()


def get_namebrand_model(model_name, num_o_classes, pretrained=False):
    known = model_name in ('inception_v3', 'alexnet', 'squeezenet') or \
        any(model_name.startswith(p) for p in ('vgg', 'resnet', 'densenet'))
    if not known:
        raise KeyError("model unknown!")
    return B200Model(model_name, num_o_classes, pretrained)


class NeustonModel(object):
    """hparams + backbone + the RUN-side hooks of the reference LightningModule."""

    def __init__(self, hparams, state_dict=None):
        if isinstance(hparams, dict):
            hparams = argparse.Namespace(**hparams)
        self.hparams = hparams
        sd = None
        if state_dict is not None:      # checkpoint keys carry the 'model.' prefix (self.model = backbone)
            sd = {k[len('model.'):]: v for k, v in state_dict.items() if k.startswith('model.')}
        self.model = B200Model(hparams.MODEL, len(hparams.classes), getattr(hparams, 'pretrained', False), sd)

    def forward(self, inputs):
        return self.model(inputs)

    def test_step(self, batch, batch_idx, dataloader_idx=None):
        input_data, input_srcs = batch
        outputs = torch.softmax(self.forward(input_data), dim=1)
        return dict(test_outputs=outputs, test_srcs=input_srcs)

    def state_dict(self):
        return {'model.' + k: v for k, v in self.model.state_dict().items()}

    # ---- checkpoint I/O (.ptl = Lightning-style torch.save dict) ----
    def save_checkpoint(self, path):
        hp = vars(self.hparams) if isinstance(self.hparams, argparse.Namespace) else dict(self.hparams)
        torch.save({'state_dict': self.state_dict(), 'hyper_parameters': hp, 'hparams_name': 'hparams',
                    'epoch': getattr(self.hparams, 'epoch', 0), 'global_step': 0,
                    'pytorch-lightning_version': '1.3.8'}, path)

    @classmethod
    def load_from_checkpoint(cls, path):
        ckpt = _load_ptl(path)
        hp = ckpt.get('hyper_parameters', ckpt.get('hparams'))
        if hp is None:
            raise KeyError('%s: no hyper_parameters in checkpoint' % path)
        if isinstance(hp, argparse.Namespace):
            hp = vars(hp)
        return cls(dict(hp), ckpt['state_dict'])

    class RunResults(object):
        def __init__(self, inputs, outputs, input_obj):
            self.inputs, self.outputs, self.input_obj = inputs, outputs, input_obj
            self.type = 'Bin' if hasattr(input_obj, 'bin_lid') else 'ImgDir'

        def __repr__(self):
            return repr('{}: {} ({} imgs)'.format(self.type, self.input_obj, len(self.inputs)))


def _load_ptl(path):
    """torch.load that tolerates references to classes of packages that are not installed
    (real Lightning checkpoints pickle callback classes as dict keys)."""
    import pickle

    class _Stub(object):
        def __init__(self, *a, **k): pass
        def __setstate__(self, state): self.__dict__['state'] = state
        def __hash__(self): return id(type(self))

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                return type(name, (_Stub,), {'__module__': module})

    class _PickleModule(object):
        Unpickler = _Unpickler
        load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
        __name__ = 'ptl_tolerant_pickle'

    return torch.load(path, map_location='cpu', weights_only=False, pickle_module=_PickleModule)
