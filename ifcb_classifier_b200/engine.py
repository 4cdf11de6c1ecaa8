"""RUN engine: classifies whole IFCB bins on one GPU.

Per bin (reference call stack neuston_net.py:253-266 -> neuston_data.py:456-464 ->
neuston_models.py:152-157 -> neuston_callbacks.py:161-162):
  raw .roi bytes + ROI table --H2D--> fused preprocess kernel (u8 gray plane)
  --> network plan (stem, tcgen05 convs, pools, head: softmax + top-1) --D2H--> scores.
Only the packed bytes cross PCIe (about 150x less than the reference's float
tensors); everything else stays in HBM.
"""
import os

import numpy as np
import torch

from . import preprocess as pp
from .graph import CompiledNet


class BinClassifier(object):
    def __init__(self, arch, state_dict, n_classes=None, img_norm=None, transform_input=False,
                 device='cuda', batch_cap=512, dtype='fp16', max_rois=4096, max_roi_bytes=64 << 20, cuda_graph=True):
        self.device = torch.device(device)
        self.net = CompiledNet(arch, state_dict, batch_cap, in_kind='u8', img_norm=img_norm,
                               transform_input=transform_input, device=self.device, dtype=dtype)
        if cuda_graph and os.environ.get('IFCB_RUN_GRAPH', '1') != '0':
            self.net.enable_cuda_graph()            # full batches replay the plan's launches from one CUDA graph
        self.R, self.batch_cap, self.n_classes = self.net.R, batch_cap, self.net.n_classes
        self._alloc(max_rois, max_roi_bytes)
        self.launches_last = 0

    def _alloc(self, max_rois, max_roi_bytes):
        d = self.device
        self.max_rois, self.max_roi_bytes = int(max_rois), int(max_roi_bytes)
        self.d_roi = torch.zeros(self.max_roi_bytes + 16, dtype=torch.uint8, device=d)
        self.d_off = torch.zeros(self.max_rois, dtype=torch.int64, device=d)
        self.d_h = torch.zeros(self.max_rois, dtype=torch.int32, device=d)
        self.d_w = torch.zeros(self.max_rois, dtype=torch.int32, device=d)
        self.d_scores = torch.zeros((self.max_rois, self.n_classes), dtype=torch.float32, device=d)
        self.d_top1 = torch.zeros(self.max_rois, dtype=torch.int32, device=d)
        self.d_top1_score = torch.zeros(self.max_rois, dtype=torch.float32, device=d)
        # pinned staging for the end-to-end path
        self.h_scores = torch.zeros((self.max_rois, self.n_classes), dtype=torch.float32).pin_memory()
        self.h_top1 = torch.zeros(self.max_rois, dtype=torch.int32).pin_memory()

    def _ensure(self, n, nbytes):
        if n > self.max_rois or nbytes > self.max_roi_bytes:
            self._alloc(max(n, self.max_rois), max(nbytes, self.max_roi_bytes))

    # ---- device-resident step -------------------------------------------------
    def upload(self, roi, offsets, heights, widths):
        """Host arrays (numpy or pinned torch tensors) -> device staging; returns (n, nbytes)."""
        n, nbytes = int(len(offsets)), int(roi.shape[0])
        self._ensure(n, nbytes)
        as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        self.d_roi[:nbytes].copy_(as_t(roi), non_blocking=True)
        self.d_off[:n].copy_(as_t(offsets), non_blocking=True)
        self.d_h[:n].copy_(as_t(heights), non_blocking=True)
        self.d_w[:n].copy_(as_t(widths), non_blocking=True)
        return n, nbytes

    def classify_device(self, n, nbytes, max_h=pp.FRAME_H, max_w=pp.FRAME_W):
        """Runs preprocess + network on the ``n`` ROIs staged by ``upload`` (async on the
        current stream).  Results land in d_scores / d_top1 / d_top1_score [:n]."""
        net, B = self.net, self.batch_cap
        launches = 0
        for i in range(0, n, B):
            m = min(B, n - i)
            pp.preprocess_rois(self.d_roi[:nbytes], self.d_off[i:i + m], self.d_h[i:i + m], self.d_w[i:i + m],
                               self.R, out_mode=pp.OUT_U8_GRAY, out=net.inp[:m], max_h=max_h, max_w=max_w)
            s, _, t1, t1s = net.forward(m)
            self.d_scores[i:i + m].copy_(s, non_blocking=True)
            self.d_top1[i:i + m].copy_(t1, non_blocking=True)
            self.d_top1_score[i:i + m].copy_(t1s, non_blocking=True)
            launches += 1 + net.num_launches
        self.launches_last = launches
        return self.d_scores[:n], self.d_top1[:n], self.d_top1_score[:n]

    # ---- end to end -------------------------------------------------------------
    def classify_bin(self, roi, offsets, heights, widths, sync=True):
        """Host bytes in -> host scores out (float32 [n, C], int32 top-1 [n])."""
        n, nbytes = self.upload(roi, offsets, heights, widths)
        if n == 0:
            return np.zeros((0, self.n_classes), np.float32), np.zeros(0, np.int32)
        mh, mw = int(np.max(np.asarray(heights))), int(np.max(np.asarray(widths)))
        s, t1, _ = self.classify_device(n, nbytes, max(mh, 1), max(mw, 1))
        self.h_scores[:n].copy_(s, non_blocking=True)
        self.h_top1[:n].copy_(t1, non_blocking=True)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return self.h_scores[:n].numpy(), self.h_top1[:n].numpy()
