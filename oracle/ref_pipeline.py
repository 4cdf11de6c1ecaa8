"""CPU restatement of the reference's per-bin RUN loop, used as the timed CPU
baseline (bench.py ``cpu_baseline`` / ``--impl reference``) and as the end-to-end
checker in tests.  TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py).

Follows the reference call stack:
  * ``IfcbBinDataset`` (/root/reference/neuston_data.py:433-467): eager ROI list,
    per item ``ToPILImage('L') -> convert('RGB') -> Resize -> ToTensor -> [Normalize]``
    -- the SAME torchvision / Pillow calls, not the numpy restatement, so the
    timing is that of the reference's own third-party code path.
  * ``DataLoader(ds, batch_size=108, pin_memory=..., num_workers=loaders)``
    (neuston_net.py:254-255; defaults :324-325).
  * ``NeustonModel.test_step`` (neuston_models.py:152-157): eval forward + softmax,
    outputs concatenated per bin (``test_epoch_end`` :159-180), then argmax / max
    (neuston_callbacks.py:161-162).
pytorch_lightning's Trainer.test wrapper (absent here) is replaced by the plain
loop it runs: model.eval(), torch.no_grad().
"""
import time

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset
from torchvision import transforms

from .pil_resize import parse_imgnorm


class RefBinDataset(Dataset):
    """Restates IfcbBinDataset over an in-memory {target: uint8[h, w]} mapping."""

    def __init__(self, images, pids, resize, img_norm=None):
        self.images, self.pids = list(images), list(pids)
        self.img_norm = parse_imgnorm(img_norm) if img_norm else None
        if isinstance(resize, int):
            resize = (resize, resize)
        self.resize = resize

    def __getitem__(self, item):
        img = self.images[item]
        img = transforms.ToPILImage(mode='L')(img)
        img = img.convert('RGB')
        img = transforms.Resize(self.resize)(img)
        img = transforms.ToTensor()(img)
        if self.img_norm:
            img = transforms.Normalize(*self.img_norm)(img)
        return img, self.pids[item]

    def __len__(self):
        return len(self.pids)


def run_bin(model, images, pids, resize, img_norm=None, batch_size=108, loaders=4):
    """One bin through the reference loop on the CPU.  Returns (scores [N,C] float32,
    classes [N], seconds dict(total, preprocess+forward overlapped))."""
    t0 = time.perf_counter()
    ds = RefBinDataset(images, pids, resize, img_norm)
    loader = DataLoader(ds, batch_size=batch_size, pin_memory=False, num_workers=loaders)
    outs = []
    model.eval()
    with torch.no_grad():
        for x, _srcs in loader:
            o = model(x)
            o = o.logits if hasattr(o, 'logits') else o
            outs.append(torch.softmax(o, dim=1))
    scores = torch.cat(outs, 0).numpy()
    dt = time.perf_counter() - t0
    return scores, np.argmax(scores, axis=1), dt
