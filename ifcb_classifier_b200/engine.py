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
        # the head writes each batch at its row of per-bin output buffers (ifcb_plan_run_at): a bin's scores are
        # contiguous on the device without any copy and leave it in one transfer
        self.window = max(1, (int(max_rois) + batch_cap - 1) // batch_cap) * batch_cap
        self.net = CompiledNet(arch, state_dict, batch_cap, in_kind='u8', img_norm=img_norm,
                               transform_input=transform_input, device=self.device, dtype=dtype, out_rows=self.window)
        if cuda_graph and os.environ.get('IFCB_RUN_GRAPH', '1') != '0':
            self.net.enable_cuda_graph()            # full batches replay the plan's launches from CUDA graphs (one per output slot)
        self.R, self.batch_cap, self.n_classes = self.net.R, batch_cap, self.net.n_classes
        self.d_scores, self.d_top1, self.d_top1_score = self.net.pb.scores, self.net.pb.top1, self.net.pb.top1_score
        self.max_roi_bytes = 0
        self._alloc_bytes(max_roi_bytes)
        d = self.device
        self.d_off = torch.zeros(self.window, dtype=torch.int64, device=d)
        self.d_h = torch.zeros(self.window, dtype=torch.int32, device=d)
        self.d_w = torch.zeros(self.window, dtype=torch.int32, device=d)
        # pinned staging for the end-to-end path (grown on demand: a bin may hold more ROIs than one window)
        self.h_scores = torch.zeros((self.window, self.n_classes), dtype=torch.float32).pin_memory()
        self.h_top1 = torch.zeros(self.window, dtype=torch.int32).pin_memory()
        self._slots = [[self.h_scores, self.h_top1, torch.cuda.Event()], None]      # two result slots: one bin in flight, one being consumed
        self._next_slot = 0
        self.launches_last = 0

    def _alloc_bytes(self, nbytes):
        if nbytes > self.max_roi_bytes:
            self.max_roi_bytes = int(nbytes)
            self.d_roi = torch.zeros(self.max_roi_bytes + 16, dtype=torch.uint8, device=self.device)

    def _ensure_host(self, n):
        if n > self.h_scores.shape[0]:
            self.h_scores = torch.zeros((n, self.n_classes), dtype=torch.float32).pin_memory()
            self.h_top1 = torch.zeros(n, dtype=torch.int32).pin_memory()
            self._slots[0] = [self.h_scores, self.h_top1, torch.cuda.Event()]

    # ---- pipelined end-to-end use (the RUN driver keeps one bin in flight while it files the previous bin's results) ----
    def submit(self, roi, offsets, heights, widths):
        """Enqueues upload + preprocess + network + download of one bin (asynchronous when ``roi`` is pinned) and returns a
        ticket for ``fetch``.  At most two tickets may be outstanding."""
        i = self._next_slot
        self._next_slot ^= 1
        n_all = int(len(offsets))
        slot = self._slots[i]
        if slot is None or slot[0].shape[0] < max(n_all, 1):
            rows = max(n_all, self.window)
            slot = self._slots[i] = [torch.zeros((rows, self.n_classes), dtype=torch.float32).pin_memory(),
                                     torch.zeros(rows, dtype=torch.int32).pin_memory(), torch.cuda.Event()]
        if i == 0:
            self.h_scores, self.h_top1 = slot[0], slot[1]
        self._classify_into(roi, offsets, heights, widths, slot[0], slot[1])
        slot[2].record(torch.cuda.current_stream(self.device))
        return i, n_all

    def fetch(self, ticket):
        """Waits for a submitted bin; returns (scores float32 [n, C], top-1 int32 [n]) -- views of the slot's pinned buffers,
        valid until the slot is reused by the second-next ``submit``."""
        i, n = ticket
        self._slots[i][2].synchronize()
        return self._slots[i][0][:n].numpy(), self._slots[i][1][:n].numpy()

    # ---- device-resident step -------------------------------------------------
    def classify_resident(self, roi, offsets, heights, widths, first=0, count=None, max_h=pp.FRAME_H, max_w=pp.FRAME_W):
        """Preprocess + network over ROIs [first, first + count) of a bin whose packed bytes and ROI table are DEVICE
        tensors (read in place, no staging copy).  count <= the output window; asynchronous on the current stream.
        Results: d_scores / d_top1 / d_top1_score [:count]."""
        n = int(offsets.shape[0]) - first if count is None else int(count)
        if n > self.window:
            raise ValueError('classify_resident: %d ROIs exceed the output window of %d (use classify_bin)' % (n, self.window))
        net, B = self.net, self.batch_cap
        launches = 0
        for i in range(0, n, B):
            m = min(B, n - i)
            lo = first + i
            pp.preprocess_rois(roi, offsets[lo:lo + m], heights[lo:lo + m], widths[lo:lo + m],
                               self.R, out_mode=pp.OUT_U8_GRAY, out=net.inp[:m], max_h=max_h, max_w=max_w)
            net.forward(m, out_row=i)
            launches += 1 + net.num_launches
        self.launches_last = launches
        return self.d_scores[:n], self.d_top1[:n], self.d_top1_score[:n]

    def upload(self, roi, offsets, heights, widths, first=0, count=None):
        """Host arrays (numpy or pinned torch tensors) -> device staging (ROI table rows [first, first+count))."""
        nbytes = int(roi.shape[0])
        n = int(len(offsets)) - first if count is None else int(count)
        as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        if first == 0:
            self._alloc_bytes(nbytes)
            self.d_roi[:nbytes].copy_(as_t(roi), non_blocking=True)
        self.d_off[:n].copy_(as_t(offsets)[first:first + n], non_blocking=True)
        self.d_h[:n].copy_(as_t(heights)[first:first + n], non_blocking=True)
        self.d_w[:n].copy_(as_t(widths)[first:first + n], non_blocking=True)
        return n, nbytes

    def classify_device(self, n, nbytes, max_h=pp.FRAME_H, max_w=pp.FRAME_W):
        """Runs preprocess + network on the ``n`` ROIs staged by ``upload`` (async on the current stream)."""
        return self.classify_resident(self.d_roi[:nbytes], self.d_off, self.d_h, self.d_w, 0, n, max_h, max_w)

    # ---- end to end -------------------------------------------------------------
    def _classify_into(self, roi, offsets, heights, widths, h_scores, h_top1):
        n_all = int(len(offsets))
        if n_all == 0:
            return
        mh, mw = max(int(np.max(np.asarray(heights))), 1), max(int(np.max(np.asarray(widths))), 1)
        launches = 0
        for w0 in range(0, n_all, self.window):                    # one pass for any bin up to `window` ROIs
            n, nbytes = self.upload(roi, offsets, heights, widths, w0, min(self.window, n_all - w0))
            s, t1, _ = self.classify_device(n, nbytes, mh, mw)
            h_scores[w0:w0 + n].copy_(s, non_blocking=True)
            h_top1[w0:w0 + n].copy_(t1, non_blocking=True)
            launches += self.launches_last
        self.launches_last = launches

    def classify_bin(self, roi, offsets, heights, widths, sync=True):
        """Host bytes in -> host scores out (float32 [n, C], int32 top-1 [n]); views of pinned buffers that the next call
        overwrites.  Pinned inputs make the upload asynchronous."""
        n_all = int(len(offsets))
        if n_all == 0:
            return np.zeros((0, self.n_classes), np.float32), np.zeros(0, np.int32)
        self._ensure_host(n_all)
        self._classify_into(roi, offsets, heights, widths, self.h_scores, self.h_top1)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return self.h_scores[:n_all].numpy(), self.h_top1[:n_all].numpy()
