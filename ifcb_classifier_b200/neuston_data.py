"""``IfcbBinDataset`` with the reference's interface (neuston_data.py:433-467), backed by
the fused preprocess kernel: the whole bin is resized / normalised in ONE launch the first
time an item is requested; items are CUDA tensors."""
import numpy as np
import torch

from . import preprocess as pp
from .preprocess import parse_imgnorm  # noqa: F401  (same name as the reference helper)


class IfcbBinDataset(object):
    def __init__(self, bin, resize, img_norm=None, device='cuda'):
        self.bin = bin
        self.img_norm = parse_imgnorm(img_norm) if img_norm else None
        if isinstance(resize, int):
            resize = (resize, resize)
        if resize[0] != resize[1]:
            raise ValueError('IfcbBinDataset: only square targets are supported (the reference uses (R, R))')
        self.resize = resize
        self.pids = list(bin.pids)
        self.device = torch.device(device)
        self._tensor = None

    @property
    def images(self):
        return [self.bin.image(i) for i in range(len(self.bin))]

    def tensor(self):
        """float32 [N, 3, R, R] on the GPU: every ROI of the bin, transformed."""
        if self._tensor is None:
            b, d = self.bin, self.device
            n = len(b)
            if n == 0:
                self._tensor = torch.zeros((0, 3) + tuple(self.resize), device=d)
            else:
                self._tensor = pp.preprocess_rois(
                    torch.from_numpy(b.roi).to(d), torch.from_numpy(b.offsets).to(d),
                    torch.from_numpy(b.heights).to(d), torch.from_numpy(b.widths).to(d), self.resize[0],
                    img_norm=self.img_norm, out_mode=pp.OUT_F32_NCHW,
                    max_h=int(b.heights.max()), max_w=int(b.widths.max()))
        return self._tensor

    def __getitem__(self, item):
        return self.tensor()[item], self.pids[item]

    def __len__(self):
        return len(self.pids)
