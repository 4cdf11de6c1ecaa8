"""CPU restatement of the reference's model construction and test step.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
``/root/reference/neuston_models.py``:
  * ``get_namebrand_model`` (lines 22-45): torchvision backbone by name with the
    classifier head swapped for ``num_o_classes`` outputs; unknown names raise
    ``KeyError("model unknown!")``.
  * ``NeustonModel.forward`` (66-68) and ``test_step`` (152-157): eval-mode
    forward, ``InceptionOutputs -> .logits`` guard, ``softmax(dim=1)``.
  * ``NeustonModel.loss`` (70-78): ``CE(out) + 0.4*CE(aux)`` for Inception.
  * ``save_run_results`` (neuston_callbacks.py:161-162): argmax / max of the scores.
The arithmetic itself is torchvision/torch (third-party, installed here:
torchvision 0.26 / torch 2.11; the reference pins 0.8.2 / 1.7.1 -- same graphs
and state_dict keys).
"""
import warnings

import numpy as np
import torch
import torch.nn as nn
import torchvision.models as MODEL_MODULE


def get_namebrand_model(model_name, num_o_classes, pretrained=False):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if model_name == 'inception_v3':
            model = MODEL_MODULE.inception_v3(pretrained)
            model.AuxLogits.fc = nn.Linear(model.AuxLogits.fc.in_features, num_o_classes)
            model.fc = nn.Linear(model.fc.in_features, num_o_classes)
        elif model_name == 'alexnet':
            model = getattr(MODEL_MODULE, model_name)(pretrained)
            model.classifier[6] = nn.Linear(model.classifier[6].in_features, num_o_classes)
        elif model_name == 'squeezenet':
            model = getattr(MODEL_MODULE, model_name + '1_1')(pretrained)
            model.classifier[1] = nn.Conv2d(512, num_o_classes, kernel_size=(1, 1), stride=(1, 1))
            model.num_classes = num_o_classes
        elif model_name.startswith('vgg'):
            model = getattr(MODEL_MODULE, model_name)(pretrained)
            model.classifier[6] = nn.Linear(model.classifier[6].in_features, num_o_classes)
        elif model_name.startswith('resnet'):
            model = getattr(MODEL_MODULE, model_name)(pretrained)
            model.fc = nn.Linear(model.fc.in_features, num_o_classes)
        elif model_name.startswith('densenet'):
            model = getattr(MODEL_MODULE, model_name)(pretrained)
            model.classifier = nn.Linear(model.classifier.in_features, num_o_classes)
        else:
            raise KeyError("model unknown!")
    return model


def test_step_scores(model, x):
    """float32 [B,3,R,R] -> softmax scores float32 [B,C] (eval, no_grad)."""
    model.eval()
    with torch.no_grad():
        out = model(x)
        out = out.logits if hasattr(out, 'logits') else out
        return torch.softmax(out, dim=1)


def loss(model_outputs, targets):
    crit = nn.CrossEntropyLoss()
    if isinstance(model_outputs, tuple) and len(model_outputs) == 2:
        return crit(model_outputs[0], targets) + 0.4 * crit(model_outputs[1], targets)
    return crit(model_outputs, targets)


def top1(scores):
    scores = np.asarray(scores)
    return np.argmax(scores, axis=1), np.max(scores, axis=1)
