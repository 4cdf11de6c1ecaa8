"""Network plans: torchvision-layout weights -> a fixed sequence of sm_100a kernel
launches over preallocated NHWC bf16 activation buffers.

Mirrors what ``get_namebrand_model`` + ``NeustonModel.forward``/``test_step``
compute in eval mode (reference neuston_models.py:22-45, 66-68, 152-157) for the
model families of BASELINE.json: ``inception_v3`` (torchvision inception.py) and
``resnet18/34/50/101/152`` (torchvision resnet.py).  torch is used here only for
device memory and one-off weight repacking at load time; every launch on the
forward path goes through the C ABI (``_lib``).

B200-first restructuring relative to the torchvision module tree:
  * BatchNorm is folded into a per-channel (scale, shift) applied in the fp32
    epilogue of the convolution kernel; ReLU and the bf16 store are fused too.
  * ``torch.cat`` disappears: every branch writes its channel slice of the block
    output in place.
  * Sibling 1x1 convolutions that read the same block input are fused into ONE
    GEMM (concatenated N) whose epilogue scatters column segments to different
    destinations -- the block input is read from HBM once instead of 3-4 times.
  * ``avg_pool2d(3,1,1)`` followed by a 1x1 conv+BN+ReLU is evaluated as
    1x1 conv (inside the fused GEMM) -> avg-pool + BN + ReLU on the conv OUTPUT
    channels (the two linear ops commute; count_include_pad=True divides by 9
    everywhere), which shrinks the pooled tensor 4-10x.
"""
import ctypes as C
import numpy as np
import torch

from . import _lib
from ._lib import (ConvDesc, StemDesc, PoolDesc, HeadDesc, IFCB_STEM_IN_U8_GRAY, IFCB_STEM_IN_F32_NCHW,
                   IFCB_POOL_MAX, IFCB_POOL_AVG_AFFINE)


class View(object):
    """Channel slice [c0, c1) of an NHWC 16-bit activation tensor.  The tensor is stored with a
    zero border of ``pad`` = (pad_h, pad_w) pixels: physical shape [B, H+2*pad_h, W+2*pad_w, C];
    H / W are the logical extent.  (The border is what the WINDOW convolution reads as padding.)"""

    def __init__(self, t, c0=0, c1=None, pad=(0, 0)):
        self.t = t
        self.c0 = c0
        self.c1 = t.shape[3] if c1 is None else c1
        self.pad = (int(pad[0]), int(pad[1]))

    @property
    def H(self): return self.t.shape[1] - 2 * self.pad[0]

    @property
    def W(self): return self.t.shape[2] - 2 * self.pad[1]

    @property
    def C(self): return self.c1 - self.c0

    @property
    def ld(self): return self.t.shape[3]

    @property
    def ptr(self): return self.t.data_ptr() + 2 * self.c0

    def slice(self, c0, c1):
        return View(self.t, self.c0 + c0, self.c0 + c1, self.pad)

    def interior(self):
        """Logical [B, H, W, C] torch view (tests / debugging)."""
        ph, pw = self.pad
        return self.t[:, ph:self.t.shape[1] - ph, pw:self.t.shape[2] - pw, self.c0:self.c1]


def fold_bn(sd, prefix, eps):
    """BatchNorm2d (eval) -> per-channel scale / shift, in float64 then float32."""
    g = sd[prefix + '.weight'].double()
    b = sd[prefix + '.bias'].double()
    m = sd[prefix + '.running_mean'].double()
    v = sd[prefix + '.running_var'].double()
    scale = g / torch.sqrt(v + eps)
    shift = b - m * scale
    return scale.float(), shift.float()


class PlanBuilder(object):
    def __init__(self, batch_cap, device, dtype='fp16', out_rows=None):
        """``dtype``: 16-bit storage / tensor-core operand format of activations and
        conv weights -- 'fp16' (default: 11-bit significand, meets the 1e-2 score
        gate) or 'bf16' (8-bit significand).  Accumulation is fp32 either way."""
        assert dtype in ('fp16', 'bf16'), dtype
        self.dtype = dtype
        self.tdtype = torch.float16 if dtype == 'fp16' else torch.bfloat16
        self.cdtype = _lib.IFCB_ACT_FP16 if dtype == 'fp16' else _lib.IFCB_ACT_BF16
        self.batch_cap = int(batch_cap)
        # rows of the head's output buffers (>= batch_cap): the batches of one bin are written side by side (run(out_row=...))
        self.out_rows = max(int(out_rows or 0), self.batch_cap)
        self.device = device
        self.keep = []                 # every device tensor the plan points into
        self.layer_names = []
        self.layer_kinds = []          # 'conv' | 'stem' | 'pool' | 'head', parallel to layer_names
        self.layer_flops = []          # algorithmic 2*MACs per image of each layer
        handle = C.c_void_p()
        _lib.check(_lib.lib().ifcb_plan_create(C.byref(handle)), 'plan_create')
        self.handle = handle
        self.flops_per_image = 0       # 2*MACs of the reference graph (algorithmic)
        self._update = None            # index into self.keep while a weight refresh replays the construction (begin_update)

    # -- weight refresh: replay the same construction sequence over the EXISTING buffers ------------------------
    def begin_update(self):
        """After this, re-running the graph builder with another state_dict overwrites the packed weights / folded BN
        vectors in place (same launches, same pointers: captured CUDA graphs stay valid); nothing is allocated or added."""
        self._update = 0

    def end_update(self):
        assert self._update == len(self.keep), 'weight refresh walked %d of %d buffers' % (self._update, len(self.keep))
        self._update = None

    def _next_kept(self):
        t = self.keep[self._update]
        self._update += 1
        return t

    def _note(self, name, kind, flops):
        if self._update is not None:
            return
        self.layer_names.append(name)
        self.layer_kinds.append(kind)
        self.layer_flops.append(flops)
        self.flops_per_image += flops

    # -- memory ---------------------------------------------------------------
    def alloc(self, H, W, Cc, pad=(0, 0)):
        """Zero-initialised activation buffer with logical extent H x W and a zero border ``pad``
        (kernels only ever write interior pixels, so the border stays zero)."""
        if self._update is not None:
            t = self._next_kept()
            assert tuple(t.shape) == (self.batch_cap, H + 2 * pad[0], W + 2 * pad[1], Cc)
            return View(t, pad=pad)
        t = torch.zeros((self.batch_cap, H + 2 * pad[0], W + 2 * pad[1], Cc), dtype=self.tdtype, device=self.device)
        self.keep.append(t)
        return View(t, pad=pad)

    def dev(self, x, dtype):
        if self._update is not None:
            t = self._next_kept()
            t.copy_(x.detach().to(dtype=dtype).reshape(t.shape))
            return t
        t = x.detach().to(device=self.device, dtype=dtype).contiguous()
        self.keep.append(t)
        return t

    @staticmethod
    def border_for(H, W, Co, k, stride=(1, 1), pad=(0, 0), Ci=64):
        """Zero border the INPUT of a conv should be allocated with: its padding when the library
        will run the conv with the WINDOW algorithm (which reads padding from memory), else none."""
        algo, _ = _lib.conv_auto_config(H, W, Ci, Co, k[0], k[1], stride, pad)
        return tuple(pad) if algo == _lib.IFCB_CONV_WINDOW else (0, 0)

    # -- layers ---------------------------------------------------------------
    def conv(self, x, members, stride=(1, 1), pad=(0, 0), residual=None, tile_n=0, name='conv', algo=0, stats=None, shift_ptr=None):
        """One implicit-GEMM launch.  ``members``: list of dicts
        (weight [Co,Ci,kh,kw] fp32, scale [Co], shift [Co], relu bool, out View) --
        more than one member = horizontally fused convs sharing input ``x``."""
        w0 = members[0]['weight']
        Ci, kh, kw = int(w0.shape[1]), int(w0.shape[2]), int(w0.shape[3])
        assert Ci == x.C, (name, Ci, x.C)
        Co = sum(int(m['weight'].shape[0]) for m in members)
        # resolve AUTO / tile_n here so that weight packing (tile_n) and the plan agree
        if algo == _lib.IFCB_CONV_AUTO:
            a_pref, _ = _lib.conv_auto_config(x.H, x.W, Ci, Co, kh, kw, stride, pad)
            can_window = tuple(stride) == (1, 1) and x.pad[0] >= pad[0] and x.pad[1] >= pad[1]
            if a_pref == _lib.IFCB_CONV_WINDOW:
                algo = _lib.IFCB_CONV_WINDOW if can_window else _lib.IFCB_CONV_IM2COL
            else:
                algo = a_pref
        if tile_n == 0:
            tile_n = _lib.lib().ifcb_conv_auto_tile_n(Co, algo)
        geo = _lib.conv_geometry(Ci, Co, kh, kw, tile_n)
        Cp, Kp, Np = geo['Cin_pad'], geo['K_pad'], geo['Cout_pad']
        wcat = torch.cat([m['weight'].float() for m in members], 0)          # [Co, Ci, kh, kw]
        packed = torch.zeros((Np, kh * kw, Cp), dtype=torch.float32)
        packed[:Co, :, :Ci] = wcat.permute(0, 2, 3, 1).reshape(Co, kh * kw, Ci)
        wdev = self.dev(packed.reshape(Np, Kp), self.tdtype)
        self.last_weight = wdev            # the packed 16-bit operand of the conv just added (TRAIN refreshes it in place)
        scale = torch.zeros(Np); shift = torch.zeros(Np)
        scale[:Co] = torch.cat([m['scale'].float() for m in members])
        shift[:Co] = torch.cat([m['shift'].float() for m in members])
        sdev, hdev = self.dev(scale, torch.float32), self.dev(shift, torch.float32)
        P = (x.H + 2 * pad[0] - kh) // stride[0] + 1
        Q = (x.W + 2 * pad[1] - kw) // stride[1] + 1
        d = ConvDesc()
        d.d_in, d.in_ld, d.Cin = x.ptr, x.ld, Ci
        d.batch_cap, d.H, d.W = self.batch_cap, x.H, x.W
        d.in_pad_h, d.in_pad_w = x.pad
        d.kh, d.kw, d.stride_h, d.stride_w, d.pad_h, d.pad_w = kh, kw, stride[0], stride[1], pad[0], pad[1]
        d.Cout = Co
        d.d_weight, d.d_scale, d.d_shift = wdev.data_ptr(), sdev.data_ptr(), hdev.data_ptr()
        if shift_ptr is not None:        # TRAIN, conv with bias: the epilogue reads the bias straight from the fp32 parameter arena
            d.d_shift = int(shift_ptr)
        d.n_seg = len(members)
        n0 = 0
        for i, m in enumerate(members):
            co = int(m['weight'].shape[0])
            out = m['out']
            assert out.C == co and out.H == P and out.W == Q, (name, out.C, co, out.H, P, out.W, Q)
            d.seg[i].n_begin, d.seg[i].n_end = n0, n0 + co
            d.seg[i].d_out, d.seg[i].ld, d.seg[i].relu = out.ptr, out.ld, 1 if m['relu'] else 0
            d.seg[i].pad_h, d.seg[i].pad_w = out.pad
            n0 += co
        if residual is not None:
            d.d_residual, d.res_ld = residual.ptr, residual.ld
            d.res_pad_h, d.res_pad_w = residual.pad
        d.tile_n = tile_n
        d.algo = algo
        d.dtype = self.cdtype
        if stats is not None:            # TRAIN: float64 [2][Co] accumulators the epilogue adds sum / sum of squares into
            d.d_stats = int(stats)
        if self._update is None:
            _lib.check(_lib.lib().ifcb_plan_add_conv(self.handle, C.byref(d)), 'plan_add_conv(%s)' % name)
        self._note(name, 'conv', 2 * P * Q * Co * Ci * kh * kw)
        return [m['out'] for m in members]

    def stem(self, inp, in_kind, H, W, weight, scale, shift, stride, pad, out, affine=None,
             in_scale=(1, 1, 1), in_shift=(0, 0, 0), name='stem', relu=True):
        """First conv.  ``affine`` = (a[3], b[3]) with x_c = a_c * g/255 + b_c for the u8 gray
        input (see ``input_affine``); f32 input uses in_scale / in_shift (transform_input)."""
        Co, _, kh, kw = [int(v) for v in weight.shape]
        d = StemDesc()
        d.d_in, d.in_kind = inp.data_ptr(), in_kind
        d.batch_cap, d.H, d.W = self.batch_cap, H, W
        d.kh, d.kw, d.stride, d.pad, d.Cout = kh, kw, stride, pad, Co
        scale, shift = scale.double(), shift.double()
        if in_kind == IFCB_STEM_IN_U8_GRAY:
            a = torch.tensor(affine[0], dtype=torch.float64)
            b = torch.tensor(affine[1], dtype=torch.float64)
            w64 = weight.double()                                               # [Co, 3, kh, kw]
            wg = (w64 * a[None, :, None, None]).sum(1) / 255.0                  # [Co, kh, kw]
            wc = (w64 * b[None, :, None, None]).sum(1)
            d.d_wgray = self.dev(wg.permute(1, 2, 0).reshape(kh * kw, Co), torch.float32).data_ptr()
            if pad > 0:
                d.d_wconst = self.dev(wc.permute(1, 2, 0).reshape(kh * kw, Co), torch.float32).data_ptr()
            else:
                shift = shift + scale * wc.sum((1, 2))                          # every tap in bounds
        else:
            wk = weight.float().permute(2, 3, 1, 0).reshape(kh * kw * 3, Co)    # k = (r*kw+s)*3 + c
            d.d_weight = self.dev(wk, torch.float32).data_ptr()
        d.d_scale = self.dev(scale, torch.float32).data_ptr()
        d.d_shift = self.dev(shift, torch.float32).data_ptr()
        for c in range(3):
            d.in_scale[c], d.in_shift[c] = float(in_scale[c]), float(in_shift[c])
        d.d_out, d.out_ld, d.relu = out.ptr, out.ld, 1 if relu else 0
        d.dtype = self.cdtype
        d.out_pad_h, d.out_pad_w = out.pad
        if self._update is None:
            _lib.check(_lib.lib().ifcb_plan_add_stem(self.handle, C.byref(d)), 'plan_add_stem')
        self._note(name, 'stem', 2 * out.H * out.W * Co * 3 * kh * kw)
        return out

    def pool(self, kind, x, k, stride, pad, out, scale=None, shift=None, relu=False, name='pool', ceil_mode=False):
        d = PoolDesc()
        d.ceil_mode = 1 if ceil_mode else 0
        d.kind, d.d_in, d.in_ld, d.C = kind, x.ptr, x.ld, x.C
        d.batch_cap, d.H, d.W, d.k, d.stride, d.pad = self.batch_cap, x.H, x.W, k, stride, pad
        d.d_out, d.out_ld = out.ptr, out.ld
        if scale is not None:
            d.d_scale = self.dev(scale, torch.float32).data_ptr()
            d.d_shift = self.dev(shift, torch.float32).data_ptr()
        d.relu = 1 if relu else 0
        d.dtype = self.cdtype
        d.in_pad_h, d.in_pad_w = x.pad
        d.out_pad_h, d.out_pad_w = out.pad
        if self._update is None:
            _lib.check(_lib.lib().ifcb_plan_add_pool(self.handle, C.byref(d)), 'plan_add_pool(%s)' % name)
        self._note(name, 'pool', 0)
        return out

    def head(self, x, weight, bias, name='head'):
        n_classes = int(weight.shape[0])
        assert x.pad == (0, 0), 'head input must be unpadded'
        B, rows = self.batch_cap, self.out_rows
        if self._update is not None:
            self.dev(weight, torch.float32)
            self.dev(bias, torch.float32)
            return self.scores
        self.scores = torch.zeros((rows, n_classes), dtype=torch.float32, device=self.device)
        self.logits = torch.zeros((rows, n_classes), dtype=torch.float32, device=self.device)
        self.top1 = torch.zeros((rows,), dtype=torch.int32, device=self.device)
        self.top1_score = torch.zeros((rows,), dtype=torch.float32, device=self.device)
        d = HeadDesc()
        d.d_in, d.in_ld, d.C, d.HW = x.ptr, x.ld, x.C, x.H * x.W
        d.batch_cap, d.n_classes = B, n_classes
        d.d_weight = self.dev(weight, torch.float32).data_ptr()
        d.d_bias = self.dev(bias, torch.float32).data_ptr()
        d.d_scores, d.d_logits = self.scores.data_ptr(), self.logits.data_ptr()
        d.d_top1, d.d_top1_score = self.top1.data_ptr(), self.top1_score.data_ptr()
        d.dtype = self.cdtype
        _lib.check(_lib.lib().ifcb_plan_add_head(self.handle, C.byref(d)), 'plan_add_head')
        self._note(name, 'head', 2 * x.C * n_classes)
        return self.scores

    def run(self, batch, first=None, last=None, out_row=0):
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if first is None:
            assert 0 <= out_row and out_row + batch <= self.out_rows, (out_row, batch, self.out_rows)
            _lib.check(_lib.lib().ifcb_plan_run_at(self.handle, int(batch), int(out_row), stream), 'plan_run')
        else:
            _lib.check(_lib.lib().ifcb_plan_run_range(self.handle, int(first), int(last), int(batch), stream),
                       'plan_run_range')

    def close(self):
        if self.handle is not None:
            _lib.lib().ifcb_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def input_lut(img_norm=None, transform_input=False):
    """float32 [3,256]: value the network sees for gray level g in channel c.

    Evaluates exactly the reference's op sequence in float32 (numpy == torch CPU):
    ToTensor ``g/255``; Normalize ``(x-mean)/std`` (neuston_data.py:462-463);
    torchvision ``_transform_input`` (inception.py:95-101) when the checkpoint
    was created with pretrained=True.
    """
    g = np.arange(256, dtype=np.float32)
    x = np.repeat((g / np.float32(255.0))[None], 3, 0)
    if img_norm:
        mean, std = img_norm
        m = np.asarray(mean, np.float32)[:, None]
        s = np.asarray(std, np.float32)[:, None]
        x = (x - m) / s
    if transform_input:
        ts, tm = transform_input_affine()
        x = x * np.asarray(ts, np.float32)[:, None] + np.asarray(tm, np.float32)[:, None]
    return torch.from_numpy(x.astype(np.float32).copy())


def input_affine(img_norm=None, transform_input=False):
    """(a[3], b[3]) in float64 with  x_c = a_c * (g/255) + b_c  -- the composition of ToTensor,
    Normalize (neuston_data.py:462-463) and torchvision's _transform_input (inception.py:95-101)."""
    a, b = [1.0, 1.0, 1.0], [0.0, 0.0, 0.0]
    if img_norm:
        mean, std = img_norm
        a = [1.0 / float(s_) for s_ in std]
        b = [-float(m_) / float(s_) for m_, s_ in zip(mean, std)]
    if transform_input:
        ts, tb = transform_input_affine()
        a = [a_ * t_ for a_, t_ in zip(a, ts)]
        b = [b_ * t_ + o_ for b_, t_, o_ in zip(b, ts, tb)]
    return a, b


def transform_input_affine():
    """(scale[3], shift[3]) of torchvision Inception3._transform_input."""
    s = [0.229 / 0.5, 0.224 / 0.5, 0.225 / 0.5]
    b = [(0.485 - 0.5) / 0.5, (0.456 - 0.5) / 0.5, (0.406 - 0.5) / 0.5]
    return s, b


# =============================================================================
# Inception-v3 (torchvision/models/inception.py), eval mode
# =============================================================================
def _basic(sd, prefix, out, relu=True, eps=1e-3):
    scale, shift = fold_bn(sd, prefix + '.bn', eps)
    return dict(weight=sd[prefix + '.conv.weight'], scale=scale, shift=shift, relu=relu, out=out)


def _raw(sd, prefix, out):
    """1x1 conv of an avg-pool branch: raw (no BN/ReLU) output, pooled afterwards."""
    w = sd[prefix + '.conv.weight']
    co = int(w.shape[0])
    return dict(weight=w, scale=torch.ones(co), shift=torch.zeros(co), relu=False, out=out)


def build_inception_v3(pb, sd, inp, in_kind, R, affine=None, transform_input=False, fuse=True):
    """Appends the eval-mode Inception-v3 graph to ``pb``; returns the softmax scores tensor."""
    eps = 1e-3
    sz = lambda h, k, s, p: (h + 2 * p - k) // s + 1

    def single(x, prefix, k, stride=(1, 1), pad=(0, 0), out=None, name=None, out_pad=(0, 0)):
        """conv+bn+relu; ``out_pad`` = zero border the NEXT conv wants around the output."""
        w = sd[prefix + '.conv.weight']
        co, kh, kw = int(w.shape[0]), int(w.shape[2]), int(w.shape[3])
        if out is None:
            out = pb.alloc(sz(x.H, kh, stride[0], pad[0]), sz(x.W, kw, stride[1], pad[1]), co, out_pad)
        pb.conv(x, [_basic(sd, prefix, out)], stride, pad, name=name or prefix)
        return out

    def fused_1x1(x, members, name):
        if fuse:
            pb.conv(x, members, name=name)
        else:
            for i, m in enumerate(members):
                pb.conv(x, [m], name='%s.%d' % (name, i))

    def avg_branch(raw, prefix, out, name):
        scale, shift = fold_bn(sd, prefix + '.bn', eps)
        pb.pool(IFCB_POOL_AVG_AFFINE, raw, 3, 1, 1, out, scale, shift, relu=True, name=name)

    # ---- stem ----
    H1 = sz(R, 3, 2, 0)
    a = pb.alloc(H1, H1, 32)
    sc, sh = fold_bn(sd, 'Conv2d_1a_3x3.bn', eps)
    ts, tb = transform_input_affine() if transform_input else ((1, 1, 1), (0, 0, 0))
    pb.stem(inp, in_kind, R, R, sd['Conv2d_1a_3x3.conv.weight'], sc, sh, 2, 0, a, affine=affine,
            in_scale=ts, in_shift=tb, name='Conv2d_1a_3x3')
    a = single(a, 'Conv2d_2a_3x3', 3, out_pad=pb.border_for(a.H - 2, a.W - 2, 64, (3, 3), pad=(1, 1)))
    a = single(a, 'Conv2d_2b_3x3', 3, pad=(1, 1))
    p = pb.alloc(sz(a.H, 3, 2, 0), sz(a.W, 3, 2, 0), 64)
    pb.pool(IFCB_POOL_MAX, a, 3, 2, 0, p, name='maxpool1')
    a = single(p, 'Conv2d_3b_1x1', 1)
    a = single(a, 'Conv2d_4a_3x3', 3)
    p = pb.alloc(sz(a.H, 3, 2, 0), sz(a.W, 3, 2, 0), 192)
    pb.pool(IFCB_POOL_MAX, a, 3, 2, 0, p, name='maxpool2')
    x = p

    # ---- InceptionA x3 ----
    for blk, pf in (('Mixed_5b', 32), ('Mixed_5c', 64), ('Mixed_5d', 64)):
        H = x.H
        out = pb.alloc(H, H, 224 + pf)
        b5, b3 = pb.border_for(H, H, 64, (5, 5), pad=(2, 2)), pb.border_for(H, H, 96, (3, 3), pad=(1, 1))
        t5, t3, tp = pb.alloc(H, H, 48, b5), pb.alloc(H, H, 64, b3), pb.alloc(H, H, pf)
        fused_1x1(x, [_basic(sd, blk + '.branch1x1', out.slice(0, 64)),
                      _basic(sd, blk + '.branch5x5_1', t5),
                      _basic(sd, blk + '.branch3x3dbl_1', t3),
                      _raw(sd, blk + '.branch_pool', tp)], blk + '.1x1s')
        single(t5, blk + '.branch5x5_2', 5, pad=(2, 2), out=out.slice(64, 128))
        t3b = single(t3, blk + '.branch3x3dbl_2', 3, pad=(1, 1), out_pad=b3)
        single(t3b, blk + '.branch3x3dbl_3', 3, pad=(1, 1), out=out.slice(128, 224))
        avg_branch(tp, blk + '.branch_pool', out.slice(224, 224 + pf), blk + '.branch_pool.avg')
        x = out

    # ---- InceptionB ----
    blk = 'Mixed_6a'
    H2 = sz(x.H, 3, 2, 0)
    out = pb.alloc(H2, H2, 768)
    single(x, blk + '.branch3x3', 3, stride=(2, 2), out=out.slice(0, 384))
    t = single(x, blk + '.branch3x3dbl_1', 1, out_pad=pb.border_for(x.H, x.W, 96, (3, 3), pad=(1, 1)))
    t = single(t, blk + '.branch3x3dbl_2', 3, pad=(1, 1))
    single(t, blk + '.branch3x3dbl_3', 3, stride=(2, 2), out=out.slice(384, 480))
    pb.pool(IFCB_POOL_MAX, x, 3, 2, 0, out.slice(480, 768), name=blk + '.maxpool')
    x = out

    # ---- InceptionC x4 ----
    for blk, c7 in (('Mixed_6b', 128), ('Mixed_6c', 160), ('Mixed_6d', 160), ('Mixed_6e', 192)):
        H = x.H
        out = pb.alloc(H, H, 768)
        bw, bh = pb.border_for(H, H, c7, (1, 7), pad=(0, 3)), pb.border_for(H, H, c7, (7, 1), pad=(3, 0))
        t7, td, tp = pb.alloc(H, H, c7, bw), pb.alloc(H, H, c7, bh), pb.alloc(H, H, 192)
        fused_1x1(x, [_basic(sd, blk + '.branch1x1', out.slice(0, 192)),
                      _basic(sd, blk + '.branch7x7_1', t7),
                      _basic(sd, blk + '.branch7x7dbl_1', td),
                      _raw(sd, blk + '.branch_pool', tp)], blk + '.1x1s')
        t = single(t7, blk + '.branch7x7_2', 7, pad=(0, 3), out_pad=bh)
        single(t, blk + '.branch7x7_3', 7, pad=(3, 0), out=out.slice(192, 384))
        t = single(td, blk + '.branch7x7dbl_2', 7, pad=(3, 0), out_pad=bw)
        t = single(t, blk + '.branch7x7dbl_3', 7, pad=(0, 3), out_pad=bh)
        t = single(t, blk + '.branch7x7dbl_4', 7, pad=(3, 0), out_pad=bw)
        single(t, blk + '.branch7x7dbl_5', 7, pad=(0, 3), out=out.slice(384, 576))
        avg_branch(tp, blk + '.branch_pool', out.slice(576, 768), blk + '.branch_pool.avg')
        x = out

    # ---- InceptionD ----
    blk = 'Mixed_7a'
    H2 = sz(x.H, 3, 2, 0)
    out = pb.alloc(H2, H2, 1280)
    bw, bh = pb.border_for(x.H, x.H, 192, (1, 7), pad=(0, 3)), pb.border_for(x.H, x.H, 192, (7, 1), pad=(3, 0))
    t3, t7 = pb.alloc(x.H, x.H, 192), pb.alloc(x.H, x.H, 192, bw)
    fused_1x1(x, [_basic(sd, blk + '.branch3x3_1', t3), _basic(sd, blk + '.branch7x7x3_1', t7)], blk + '.1x1s')
    single(t3, blk + '.branch3x3_2', 3, stride=(2, 2), out=out.slice(0, 320))
    t = single(t7, blk + '.branch7x7x3_2', 7, pad=(0, 3), out_pad=bh)
    t = single(t, blk + '.branch7x7x3_3', 7, pad=(3, 0))
    single(t, blk + '.branch7x7x3_4', 3, stride=(2, 2), out=out.slice(320, 512))
    pb.pool(IFCB_POOL_MAX, x, 3, 2, 0, out.slice(512, 1280), name=blk + '.maxpool')
    x = out

    # ---- InceptionE x2 ----
    for blk in ('Mixed_7b', 'Mixed_7c'):
        H = x.H
        out = pb.alloc(H, H, 2048)
        b3 = pb.border_for(H, H, 384, (3, 3), pad=(1, 1))       # (1x3 / 3x1 siblings share the t3 buffer)
        t3, td, tp = pb.alloc(H, H, 384, b3), pb.alloc(H, H, 448, b3), pb.alloc(H, H, 192)
        fused_1x1(x, [_basic(sd, blk + '.branch1x1', out.slice(0, 320)),
                      _basic(sd, blk + '.branch3x3_1', t3),
                      _basic(sd, blk + '.branch3x3dbl_1', td),
                      _raw(sd, blk + '.branch_pool', tp)], blk + '.1x1s')
        single(t3, blk + '.branch3x3_2a', 3, pad=(0, 1), out=out.slice(320, 704))
        single(t3, blk + '.branch3x3_2b', 3, pad=(1, 0), out=out.slice(704, 1088))
        t = single(td, blk + '.branch3x3dbl_2', 3, pad=(1, 1), out_pad=b3)
        single(t, blk + '.branch3x3dbl_3a', 3, pad=(0, 1), out=out.slice(1088, 1472))
        single(t, blk + '.branch3x3dbl_3b', 3, pad=(1, 0), out=out.slice(1472, 1856))
        avg_branch(tp, blk + '.branch_pool', out.slice(1856, 2048), blk + '.branch_pool.avg')
        x = out

    return pb.head(x, sd['fc.weight'], sd['fc.bias'])


# =============================================================================
# ResNet (torchvision/models/resnet.py), eval mode
# =============================================================================
RESNET_CFG = {
    'resnet18': ('basic', [2, 2, 2, 2]), 'resnet34': ('basic', [3, 4, 6, 3]),
    'resnet50': ('bottleneck', [3, 4, 6, 3]), 'resnet101': ('bottleneck', [3, 4, 23, 3]),
    'resnet152': ('bottleneck', [3, 8, 36, 3]),
}


def build_resnet(pb, sd, arch, inp, in_kind, R, affine=None):
    eps = 1e-5
    kind, layers = RESNET_CFG[arch]
    sz = lambda h, k, s, p: (h + 2 * p - k) // s + 1

    def cbr(x, conv, bn, stride=1, pad=0, relu=True, residual=None, name=None, out_pad=(0, 0)):
        w = sd[conv + '.weight']
        co, k = int(w.shape[0]), int(w.shape[2])
        out = pb.alloc(sz(x.H, k, stride, pad), sz(x.W, k, stride, pad), co, out_pad)
        scale, shift = fold_bn(sd, bn, eps)
        pb.conv(x, [dict(weight=w, scale=scale, shift=shift, relu=relu, out=out)], (stride, stride), (pad, pad),
                residual=residual, name=name or conv)
        return out

    # a tensor carries the zero border its 3x3 consumer needs (basic blocks: block inputs/outputs)
    H1 = sz(R, 7, 2, 3)
    a = pb.alloc(H1, H1, 64)
    sc, sh = fold_bn(sd, 'bn1', eps)
    pb.stem(inp, in_kind, R, R, sd['conv1.weight'], sc, sh, 2, 3, a, affine=affine, name='conv1')
    H2 = sz(H1, 3, 2, 1)
    width0 = 64
    x = pb.alloc(H2, H2, 64, pb.border_for(H2, H2, width0, (3, 3), pad=(1, 1)) if kind == 'basic' else (0, 0))
    pb.pool(IFCB_POOL_MAX, a, 3, 2, 1, x, name='maxpool')
    # a tensor carries the zero border its 3x3 consumer wants (only when that conv will run the
    # WINDOW algorithm: PlanBuilder.border_for)
    blocks = [(li, bi) for li, nb in enumerate(layers) for bi in range(nb)]
    for idx, (li, bi) in enumerate(blocks):
        pre = 'layer%d.%d' % (li + 1, bi)
        stride = 2 if (li > 0 and bi == 0) else 1
        width = 64 << li
        Ho = sz(x.H, 3, stride, 1)
        identity = x
        if (pre + '.downsample.0.weight') in sd:
            identity = cbr(x, pre + '.downsample.0', pre + '.downsample.1', stride=stride, relu=False)
        if kind == 'basic':
            # border of the block output = what the NEXT block's conv1 (3x3, maybe stride 2) wants
            if idx + 1 < len(blocks):
                nli, nbi = blocks[idx + 1]
                nstride = 2 if (nli > 0 and nbi == 0) else 1
                opad = pb.border_for(Ho, Ho, 64 << nli, (3, 3), (nstride, nstride), (1, 1))
            else:
                opad = (0, 0)
            t = cbr(x, pre + '.conv1', pre + '.bn1', stride=stride, pad=1,
                    out_pad=pb.border_for(Ho, Ho, width, (3, 3), pad=(1, 1)))
            x = cbr(t, pre + '.conv2', pre + '.bn2', pad=1, relu=True, residual=identity, out_pad=opad)
        else:   # torchvision v1.5: the stride sits on the 3x3
            t = cbr(x, pre + '.conv1', pre + '.bn1',
                    out_pad=pb.border_for(x.H, x.W, width, (3, 3), (stride, stride), (1, 1)))
            t = cbr(t, pre + '.conv2', pre + '.bn2', stride=stride, pad=1)
            x = cbr(t, pre + '.conv3', pre + '.bn3', relu=True, residual=identity)
    return pb.head(x, sd['fc.weight'], sd['fc.bias'])


# =============================================================================
# VGG / AlexNet (torchvision/models/vgg.py, alexnet.py), eval mode: conv(+bias)(+BN)+ReLU stacks, max pools,
# and the three-layer classifier run as convolutions on the same tcgen05 kernel (Linear over the
# flattened 7x7 / 6x6 feature map == a 7x7 / 6x6 conv to a 1x1 output; Dropout is the identity in eval).
# =============================================================================
VGG_CFG = {
    'vgg11': [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'vgg13': [64, 64, 'M', 128, 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512, 'M'],
    'vgg16': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],
    'vgg19': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M'],
}
# AlexNet features: (index in nn.Sequential, kernel, stride, pad) / 'M' = MaxPool2d(3, 2)
ALEXNET_CFG = [(0, 11, 4, 2), 'M', (3, 5, 1, 2), 'M', (6, 3, 1, 1), (8, 3, 1, 1), (10, 3, 1, 1), 'M']
PLAIN_ARCHS = tuple(VGG_CFG) + tuple(a + '_bn' for a in VGG_CFG) + ('alexnet',)


def build_plain_cnn(pb, sd, arch, inp, in_kind, R, affine=None):
    """VGG (with or without BatchNorm) and AlexNet."""
    sz = lambda h, k, s, p: (h + 2 * p - k) // s + 1

    def layer(idx, bn_idx):
        w, b = sd['features.%d.weight' % idx], sd['features.%d.bias' % idx].float()
        if bn_idx is not None:                                   # conv bias folds into the BN shift
            scale, shift = fold_bn(sd, 'features.%d' % bn_idx, 1e-5)
            return w, scale, shift + scale * b
        return w, torch.ones(int(w.shape[0])), b

    if arch == 'alexnet':
        seq = list(ALEXNET_CFG)
        pool_k, pool_s = 3, 2
    else:
        bn = arch.endswith('_bn')
        seq, idx = [], 0
        for v in VGG_CFG[arch[:-3] if bn else arch]:
            if v == 'M':
                seq.append('M'); idx += 1
            else:
                seq.append((idx, 3, 1, 1, idx + 1 if bn else None)); idx += 3 if bn else 2
        pool_k, pool_s = 2, 2
    x, H, first = None, R, True
    for pos, item in enumerate(seq):
        if item == 'M':
            Ho = sz(H, pool_k, pool_s, 0)
            out = pb.alloc(Ho, Ho, x.C)
            pb.pool(IFCB_POOL_MAX, x, pool_k, pool_s, 0, out, name='features.pool%d' % pos)
            x, H = out, Ho
            continue
        idx, k, stride, pad = item[:4]
        w, scale, shift = layer(idx, item[4] if len(item) > 4 else None)
        Ho = sz(H, k, stride, pad)
        nxt = seq[pos + 1] if pos + 1 < len(seq) else 'M'
        opad = pb.border_for(Ho, Ho, int(w.shape[0]), (nxt[1], nxt[1]), (nxt[2], nxt[2]), (nxt[3], nxt[3]), Ci=int(w.shape[0])) if nxt != 'M' else (0, 0)
        out = pb.alloc(Ho, Ho, int(w.shape[0]), opad)
        if first:
            pb.stem(inp, in_kind, R, R, w, scale, shift, stride, pad, out, affine=affine, name='features.%d' % idx)
            first = False
        else:
            pb.conv(x, [dict(weight=w, scale=scale, shift=shift, relu=True, out=out)], (stride, stride), (pad, pad), name='features.%d' % idx)
        x, H = out, Ho
    # classifier: [Dropout] Linear ReLU [Dropout] Linear ReLU Linear  (AdaptiveAvgPool2d is the identity at 224 px)
    lin = sorted(int(k.split('.')[1]) for k in sd if k.startswith('classifier.') and k.endswith('.weight'))
    assert len(lin) == 3
    feat = int(sd['classifier.%d.weight' % lin[0]].shape[1])
    assert feat == x.C * H * H, '%s: input size %d gives a %dx%d feature map; the classifier expects %d features (use %d px)' % (
        arch, R, H, H, feat, 224)
    w1 = sd['classifier.%d.weight' % lin[0]].view(-1, x.C, H, H)                      # flatten order of NCHW: (c, h, w)
    h1 = pb.alloc(1, 1, int(w1.shape[0]))
    pb.conv(x, [dict(weight=w1, scale=torch.ones(int(w1.shape[0])), shift=sd['classifier.%d.bias' % lin[0]].float(), relu=True, out=h1)],
            name='classifier.%d' % lin[0])
    w2 = sd['classifier.%d.weight' % lin[1]]
    h2 = pb.alloc(1, 1, int(w2.shape[0]))
    pb.conv(h1, [dict(weight=w2.view(w2.shape[0], w2.shape[1], 1, 1), scale=torch.ones(int(w2.shape[0])),
                      shift=sd['classifier.%d.bias' % lin[1]].float(), relu=True, out=h2)], name='classifier.%d' % lin[1])
    return pb.head(h2, sd['classifier.%d.weight' % lin[2]], sd['classifier.%d.bias' % lin[2]])


# =============================================================================
# SqueezeNet 1.1 (torchvision/models/squeezenet.py; the reference's 'squeezenet' = squeezenet1_1 with a
# 1x1-conv classifier of n_classes outputs, neuston_models.py:30-33), eval mode
# =============================================================================
def build_squeezenet(pb, sd, inp, in_kind, R, affine=None):
    sz = lambda h, k, s, p: (h + 2 * p - k) // s + 1

    def ceil_sz(h, k, s):
        o = (h - k + s - 1) // s + 1
        return o - 1 if (o - 1) * s >= h else o

    def cbr(prefix, out):
        w = sd[prefix + '.weight']
        return dict(weight=w, scale=torch.ones(int(w.shape[0])), shift=sd[prefix + '.bias'].float(), relu=True, out=out)

    H = sz(R, 3, 2, 0)
    x = pb.alloc(H, H, 64)
    w0 = sd['features.0.weight']
    pb.stem(inp, in_kind, R, R, w0, torch.ones(64), sd['features.0.bias'].float(), 2, 0, x, affine=affine, name='features.0')
    # features: 0 conv, 1 relu, 2 pool, 3-4 fire, 5 pool, 6-7 fire, 8 pool, 9-12 fire
    for idx in (2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12):
        if idx in (2, 5, 8):
            Ho = ceil_sz(x.H, 3, 2)
            out = pb.alloc(Ho, Ho, x.C)
            pb.pool(IFCB_POOL_MAX, x, 3, 2, 0, out, name='features.%d' % idx, ceil_mode=True)
            x = out
            continue
        pre = 'features.%d' % idx
        sq = int(sd[pre + '.squeeze.weight'].shape[0])
        e1, e3 = int(sd[pre + '.expand1x1.weight'].shape[0]), int(sd[pre + '.expand3x3.weight'].shape[0])
        t = pb.alloc(x.H, x.W, sq, pb.border_for(x.H, x.W, e3, (3, 3), pad=(1, 1), Ci=sq))
        pb.conv(x, [cbr(pre + '.squeeze', t)], name=pre + '.squeeze')
        out = pb.alloc(x.H, x.W, e1 + e3)
        pb.conv(t, [cbr(pre + '.expand1x1', out.slice(0, e1))], name=pre + '.expand1x1')
        pb.conv(t, [cbr(pre + '.expand3x3', out.slice(e1, e1 + e3))], pad=(1, 1), name=pre + '.expand3x3')
        x = out
    # classifier: Dropout, Conv2d(512, n_classes, 1), ReLU, AdaptiveAvgPool2d(1)  ->  softmax over the pooled map
    wc, bc = sd['classifier.1.weight'], sd['classifier.1.bias'].float()
    n_classes = int(wc.shape[0])
    c16 = (n_classes + 15) // 16 * 16
    wp = torch.zeros((c16, int(wc.shape[1]), 1, 1)); wp[:n_classes] = wc
    bp_ = torch.zeros(c16); bp_[:n_classes] = bc
    y = pb.alloc(x.H, x.W, c16)
    pb.conv(x, [dict(weight=wp, scale=torch.ones(c16), shift=bp_, relu=True, out=y)], name='classifier.1')
    eye = torch.zeros((n_classes, c16)); eye[:, :n_classes] = torch.eye(n_classes)       # head = mean over pixels -> identity "fc" -> softmax
    return pb.head(y, eye, torch.zeros(n_classes))


# =============================================================================
# DenseNet-121/169/201 (torchvision/models/densenet.py), eval mode.  Pre-activation layers: every dense layer has its
# own BatchNorm over ALL features so far, so norm1 + relu1 cannot fold into a producer's epilogue -- it runs as a
# k = 1 affine "pool" launch into a scratch tensor; norm2 + relu2 fold into conv1's epilogue; conv2 writes its growth
# channels straight into the block's feature buffer (the concat is in place).
# =============================================================================
def build_densenet(pb, sd, inp, in_kind, R, affine=None):
    eps = 1e-5
    sz = lambda h, k, s, p: (h + 2 * p - k) // s + 1
    n_init = int(sd['features.conv0.weight'].shape[0])
    assert n_init in (32, 64, 96), n_init
    blocks = []
    b = 1
    while ('features.denseblock%d.denselayer1.conv1.weight' % b) in sd:
        n = 1
        while ('features.denseblock%d.denselayer%d.conv1.weight' % (b, n + 1)) in sd:
            n += 1
        blocks.append(n)
        b += 1
    growth = int(sd['features.denseblock1.denselayer1.conv2.weight'].shape[0])
    H0 = sz(R, 7, 2, 3)
    a = pb.alloc(H0, H0, n_init)
    sc, sh = fold_bn(sd, 'features.norm0', eps)
    pb.stem(inp, in_kind, R, R, sd['features.conv0.weight'], sc, sh, 2, 3, a, affine=affine, name='features.conv0')
    H = sz(H0, 3, 2, 1)
    C_in = n_init
    buf = pb.alloc(H, H, C_in + blocks[0] * growth)
    pb.pool(IFCB_POOL_MAX, a, 3, 2, 1, buf.slice(0, C_in), name='features.pool0')
    for bi, n_layers in enumerate(blocks):
        for li in range(n_layers):
            pre = 'features.denseblock%d.denselayer%d' % (bi + 1, li + 1)
            xin = buf.slice(0, C_in)
            t1 = pb.alloc(H, H, C_in)
            s1, h1 = fold_bn(sd, pre + '.norm1', eps)
            pb.pool(IFCB_POOL_AVG_AFFINE, xin, 1, 1, 0, t1, s1, h1, relu=True, name=pre + '.norm1')
            w1 = sd[pre + '.conv1.weight']
            mid = int(w1.shape[0])
            t2 = pb.alloc(H, H, mid, pb.border_for(H, H, growth, (3, 3), pad=(1, 1), Ci=mid))
            s2, h2 = fold_bn(sd, pre + '.norm2', eps)
            pb.conv(t1, [dict(weight=w1, scale=s2, shift=h2, relu=True, out=t2)], name=pre + '.conv1')
            w2 = sd[pre + '.conv2.weight']
            pb.conv(t2, [dict(weight=w2, scale=torch.ones(growth), shift=torch.zeros(growth), relu=False,
                              out=buf.slice(C_in, C_in + growth))], pad=(1, 1), name=pre + '.conv2')
            C_in += growth
        if bi + 1 < len(blocks):                       # transition: norm -> relu -> conv 1x1 -> avg_pool 2x2
            pre = 'features.transition%d' % (bi + 1)
            t1 = pb.alloc(H, H, C_in)
            s1, h1 = fold_bn(sd, pre + '.norm', eps)
            pb.pool(IFCB_POOL_AVG_AFFINE, buf, 1, 1, 0, t1, s1, h1, relu=True, name=pre + '.norm')
            w = sd[pre + '.conv.weight']
            Co = int(w.shape[0])
            t2 = pb.alloc(H, H, Co)
            pb.conv(t1, [dict(weight=w, scale=torch.ones(Co), shift=torch.zeros(Co), relu=False, out=t2)], name=pre + '.conv')
            H = sz(H, 2, 2, 0)
            nbuf = pb.alloc(H, H, Co + blocks[bi + 1] * growth)
            pb.pool(IFCB_POOL_AVG_AFFINE, t2, 2, 2, 0, nbuf.slice(0, Co), torch.ones(Co), torch.zeros(Co), relu=False, name=pre + '.pool')
            buf, C_in = nbuf, Co
    t = pb.alloc(H, H, C_in)
    s5, h5 = fold_bn(sd, 'features.norm5', eps)
    pb.pool(IFCB_POOL_AVG_AFFINE, buf, 1, 1, 0, t, s5, h5, relu=True, name='features.norm5')
    return pb.head(t, sd['classifier.weight'], sd['classifier.bias'])


DENSE_ARCHS = ('densenet121', 'densenet161', 'densenet169', 'densenet201')

class CompiledNet(object):
    """A model compiled for a fixed batch capacity and input kind.

    ``in_kind`` 'u8': input is the resized gray plane uint8 [B, R, R] produced by
    the preprocess kernel (ToTensor/Normalize/transform_input folded into a LUT);
    'f32': input is the reference's float32 [B, 3, R, R] tensor (drop-in forward).
    """

    def __init__(self, arch, state_dict, batch_cap, in_kind='u8', R=None, img_norm=None,
                 transform_input=False, device='cuda', fuse=True, dtype='fp16', out_rows=None):
        self.arch = arch
        self.R = R or (299 if arch == 'inception_v3' else 224)
        self.in_kind = in_kind
        self.batch_cap = int(batch_cap)
        self.device = torch.device(device)
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        pb = PlanBuilder(batch_cap, self.device, dtype, out_rows=out_rows)
        self.out_rows = pb.out_rows
        self.dtype = dtype
        if in_kind == 'u8':
            self.inp = torch.zeros((batch_cap, self.R, self.R), dtype=torch.uint8, device=self.device)
            kind = IFCB_STEM_IN_U8_GRAY
            affine = input_affine(img_norm, transform_input and arch == 'inception_v3')
        else:
            self.inp = torch.zeros((batch_cap, 3, self.R, self.R), dtype=torch.float32, device=self.device)
            kind = IFCB_STEM_IN_F32_NCHW
            affine = None
        self._build_args = dict(kind=kind, affine=affine, transform_input=transform_input, fuse=fuse)
        self._build(pb, sd)
        self.pb = pb
        self.n_classes = int(pb.scores.shape[1])
        self.flops_per_image = pb.flops_per_image
        self.num_launches = _lib.lib().ifcb_plan_num_launches(pb.handle)

    def _build(self, pb, sd):
        arch, a = self.arch, self._build_args
        if arch == 'inception_v3':
            build_inception_v3(pb, sd, self.inp, a['kind'], self.R, affine=a['affine'], transform_input=a['transform_input'], fuse=a['fuse'])
        elif arch in RESNET_CFG:
            build_resnet(pb, sd, arch, self.inp, a['kind'], self.R, affine=a['affine'])
        elif arch in PLAIN_ARCHS:
            build_plain_cnn(pb, sd, arch, self.inp, a['kind'], self.R, affine=a['affine'])
        elif arch == 'squeezenet':
            build_squeezenet(pb, sd, self.inp, a['kind'], self.R, affine=a['affine'])
        elif arch.startswith('densenet'):
            build_densenet(pb, sd, self.inp, a['kind'], self.R, affine=a['affine'])
        else:
            raise KeyError('model unknown!')

    def load_state_dict(self, state_dict):
        """Refreshes the packed weights and folded BatchNorm vectors IN PLACE from another state_dict of the same architecture
        (TRAIN validates every epoch with the current weights: no new plan, no new activation buffers, graphs stay valid)."""
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.pb.begin_update()
        self._build(self.pb, sd)
        self.pb.end_update()
        torch.cuda.synchronize(self.device)
        _lib.check(_lib.lib().ifcb_plan_refresh(self.pb.handle), 'plan_refresh')
        if getattr(self, '_graphs', None):          # the stem's constant-bank weights are kernel PARAMETERS baked into captured graphs
            self.enable_cuda_graph()

    def enable_cuda_graph(self):
        """Captures the full-batch forward (every launch of the plan) into CUDA graphs, one per output slot
        (``out_row`` = 0, batch_cap, 2*batch_cap, ... < out_rows); ``forward(batch_cap, out_row)`` then replays the slot's
        graph.  All pointers of a plan are fixed at construction, so the graphs stay valid."""
        for _ in range(2):
            self.pb.run(self.batch_cap)
        torch.cuda.synchronize(self.device)
        self._graphs = {}
        for row in range(0, self.out_rows - self.batch_cap + 1, self.batch_cap):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.pb.run(self.batch_cap, out_row=row)
            self._graphs[row] = g

    def forward(self, n, out_row=0):
        """Runs the plan on the first ``n`` images of ``self.inp`` (current stream); the head writes rows
        [out_row, out_row + n) of the output buffers.  Returns views (scores, logits, top1, top1_score) of those rows."""
        g = getattr(self, '_graphs', {}).get(out_row) if n == self.batch_cap else None
        if g is not None:
            g.replay()
        else:
            self.pb.run(n, out_row=out_row)
        sl = slice(out_row, out_row + n)
        return self.pb.scores[sl], self.pb.logits[sl], self.pb.top1[sl], self.pb.top1_score[sl]
