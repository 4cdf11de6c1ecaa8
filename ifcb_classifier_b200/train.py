"""TRAIN step on the B200 kernels: train-mode forward, backward and Adam for the torchvision
graphs ``get_namebrand_model`` builds (reference neuston_models.py:22-45) as
``NeustonModel.training_step`` / ``loss`` / ``configure_optimizers`` run them
(neuston_models.py:63-86): CrossEntropyLoss (+ 0.4 * aux loss for Inception-v3), Adam(lr=1e-3),
per-rank BatchNorm statistics, gradient mean over ranks (Lightning DDP, neuston_net.py:101-107).

torch supplies device memory, streams and ``torch.distributed``; every kernel on the step goes
through the C ABI (``_lib``): tcgen05 implicit-GEMM convolutions for the forward and the data
gradient (the same kernel, fed the flipped/transposed filter), the tcgen05 weight-gradient kernel,
and the streaming kernels of ``csrc/train_ops.cu``.

Layout.  Parameters live in ONE flat fp32 arena (``params``) with a parallel gradient arena and
the two Adam moment arenas, so that the optimizer is a single launch and the gradient exchange a
handful of NCCL all-reduces over contiguous buckets launched while the backward pass is still
running.  Conv master weights are stored ``[Cout, kh*kw, Cin]`` (the layout the weight-gradient
kernel accumulates into); ``state_dict()`` converts back to torchvision's ``[Cout, Cin, kh, kw]``.
Activations are NHWC 16-bit (bf16 by default for training); for each conv+BN unit both the conv
output ``z`` and the activation ``a`` are kept for the backward pass.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ViewDesc, WgradDesc
from .graph import PlanBuilder, View, RESNET_CFG, PLAIN_ARCHS, VGG_CFG, ALEXNET_CFG
from .sharding import plan_buckets, GradReducer


def _vd(v):
    d = ViewDesc()
    d.d, d.ld, d.C, d.H, d.W = v.ptr, v.ld, v.C, v.H, v.W
    d.pad_h, d.pad_w = v.pad
    return d


def build_dgrad(bp, dy, dx, Co, Ci, kh, kw, stride, pad, accumulate, name='dgrad'):
    """Appends the data-gradient of a Conv2d(Ci -> Co, kh x kw, stride, pad) to plan ``bp``:
    dx (=|+=) conv_transpose(dy, W).  The transposed conv is the forward tcgen05 kernel run at stride 1
    over the (zero-dilated, for stride > 1) output gradient with padding k-1-pad and the operand
    ``ifcb_conv_repack`` writes (taps reversed, Cin/Cout swapped).  ``dy`` / ``dx``: gradient Views
    (``dy`` may carry the zero border k-1-pad, which lets the WINDOW scheme run the transposed conv).  Returns dict(run=[closures], weight=<16-bit operand tensor>, Cin_pad=<its channel padding>)."""
    B, lib = bp.batch_cap, _lib.lib()
    run = []
    stream = lambda: C.c_void_p(torch.cuda.current_stream(bp.device).cuda_stream)
    dpad = (kh - 1 - pad[0], kw - 1 - pad[1])                     # padding of the transposed conv
    if tuple(stride) != (1, 1):
        Hd, Wd = dx.H + 2 * pad[0] - kh + 1, dx.W + 2 * pad[1] - kw + 1
        bh, bw = bp.border_for(Hd, Wd, Ci, (kh, kw), pad=dpad, Ci=Co)       # zero border if the WINDOW scheme will run it
        src = View(torch.zeros((B, Hd + 2 * bh, Wd + 2 * bw, Co), dtype=bp.tdtype, device=bp.device), pad=(bh, bw))
        bp.keep.append(src.t)
        sd_, dd_ = _vd(dy), _vd(src)
        run.append(lambda: _lib.check(lib.ifcb_dilate(C.byref(sd_), C.byref(dd_), B, stride[0], stride[1], stream()), 'dilate'))
    else:
        src = dy
    li = len(bp.layer_names)
    wflip = torch.zeros((Ci, Co, kh, kw))                           # placeholder: ifcb_conv_repack fills the operand
    bp.conv(src, [dict(weight=wflip, scale=torch.ones(Ci), shift=torch.zeros(Ci), relu=False, out=dx)], (1, 1), dpad,
            residual=dx if accumulate else None, name=name)
    wdg = bp.last_weight
    cin_pad = _lib.conv_geometry(Co, Ci, kh, kw)['Cin_pad']
    assert wdg.shape[1] == kh * kw * cin_pad, wdg.shape
    run.append(lambda: bp.run(B, li, li + 1))
    return dict(run=run, weight=wdg, Cin_pad=cin_pad)


class _Param(object):
    """A slice of the flat parameter / gradient arenas."""

    def __init__(self, net, name, off, shape, kind, meta=None):
        self.net, self.name, self.off, self.shape, self.kind, self.meta = net, name, off, tuple(shape), kind, meta or {}
        self.n = 1
        for s in shape:
            self.n *= int(s)

    @property
    def w(self): return self.net.params[self.off:self.off + self.n].view(self.shape)

    @property
    def g(self): return self.net.grads[self.off:self.off + self.n].view(self.shape)

    @property
    def wptr(self): return self.net.params.data_ptr() + 4 * self.off

    @property
    def gptr(self): return self.net.grads.data_ptr() + 4 * self.off


class TrainNet(object):
    """``arch`` in train mode for a fixed per-GPU batch.

    ``step(x, labels)`` = one ``training_step`` + ``backward`` + ``optimizer.step``:
    x float32 [B,3,R,R] (what the reference's DataLoader yields), labels int64 [B]; returns the
    loss as a 0-dim device tensor (no host sync).  ``forward_backward`` stops before Adam.
    """

    def __init__(self, arch, state_dict, batch, device='cuda', dtype='bf16', lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 dropout=True, seed=0, R=None, bucket_mb=32, keep_dy=False, window=True, transform_input=False, share=None,
                 deterministic=False):
        """``share``: another TrainNet of the same architecture whose parameter / gradient / Adam arenas, BatchNorm running
        statistics and step counter this one uses instead of allocating its own -- a second plan for another batch size over
        the SAME model (the short last batch of an epoch, neuston_models.py:80-86 trains on it as is).  The two plans keep
        separate 16-bit operand copies: call ``repack()`` on the one about to step after the other has stepped."""
        self.arch, self.batch, self.device = arch, int(batch), torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('TrainNet: a CUDA device is required (there is no CPU path)')
        self.R = R or (299 if arch == 'inception_v3' else 224)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.dropout, self.seed = bool(dropout), int(seed)
        # zero borders on activations / gradients so that k > 1 stride-1 convs run the WINDOW scheme (IFCB_TRAIN_WINDOW=0: A/B switch)
        self.window = bool(window) and os.environ.get('IFCB_TRAIN_WINDOW', '1') != '0'
        # torchvision Inception3.transform_input (set by the factory when pretrained weights are requested, inception.py:95-101)
        self.transform_input = bool(transform_input) and arch == 'inception_v3'
        self.keep_dy = bool(keep_dy)        # tests: keep d(activation) next to d(conv output) instead of overwriting it
        self._share = share
        # Trainer(deterministic=True) upstream (neuston_net.py:101): two-pass split-K weight gradients and ordered BatchNorm
        # reductions instead of floating-point atomics (ifcb_train_deterministic): bitwise reproducible steps, a few % slower
        self.deterministic = bool(deterministic)
        self._det_need = 2 << 20
        self._steps = share._steps if share is not None else [0]      # optimizer step counter (shared between plans of one model)
        self._reducer = None
        self.lib = _lib.lib()
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.sd_keys = list(sd.keys())
        self.fp = PlanBuilder(batch, self.device, dtype)       # forward convs (+ stem)
        self.bp = PlanBuilder(batch, self.device, dtype)       # data-gradient convs
        self.cdtype, self.tdtype = self.fp.cdtype, self.fp.tdtype
        total = sum(((v.numel() * (3 if v.dim() == 4 and v.shape[1] == 3 else 1) + 63) // 64) * 64 + 64
                    for v in sd.values() if v.is_floating_point()) + (1 << 20)
        self.params = share.params if share is not None else torch.zeros(total, dtype=torch.float32, device=self.device)
        self._cursor = 0
        self.plist = []                 # _Param in forward order
        self.buffers = {}               # running_mean / running_var / num_batches_tracked
        self.records = []               # forward records, replayed backwards by _finalize
        self.pre = []                   # closures that must run eagerly before the forward pass (step-dependent kernel arguments)
        self.fwd = []                   # closures
        self._nbt_keys = []             # BatchNorm num_batches_tracked buffers (host counters)
        self._graph = None
        self._repack_table = None
        self.bwd = []                   # closures, already in execution order
        self.repacks = []               # (param, Co, taps, Ci, forward operand, Cin_pad, data-gradient operand or None, Cout_padk)
        self.keep = []
        self._grad_t = {}               # id(activation tensor) -> gradient tensor
        self._grad_pad = {}             # id(activation tensor) -> zero border of its gradient tensor
        self.inp = torch.zeros((batch, 3, self.R, self.R), dtype=torch.float32, device=self.device)
        self.labels = torch.zeros((batch,), dtype=torch.int64, device=self.device)
        self.loss = torch.zeros((2,), dtype=torch.float32, device=self.device)     # [main + aux (weighted), unused]
        self.acc = torch.zeros((8192,), dtype=torch.float64, device=self.device)      # 64 KB BN scratch: float64 sums [0,4096), coefficient floats behind
        # OPTIONAL: BatchNorm batch statistics gathered by the conv epilogues (ifcb_conv_desc.d_stats -> ifcb_bn_apply_sums): float64
        # [2][Co] per unit, zeroed once per step, no statistics pass over z.  Parity-tested, but OFF by default: the ~60 extra
        # epilogue instructions per 32x32 unit cost more than the removed pass on this kernel (the training convs are epilogue /
        # HBM bound, not MMA bound) -- measured at batch 256: ResNet-50 9.46 k img/s without, 9.05 k with it on every conv, 9.31 k
        # with it on K >= 512 convs only; Inception-v3 7.31 / 7.06 / 7.21 k (profiles/r02_train_epilogue_stats_ab.txt).
        # IFCB_TRAIN_EPI_STATS=1 switches it on for convs with K = kh*kw*Cin >= IFCB_TRAIN_EPI_STATS_MINK.
        self.epi_stats = os.environ.get('IFCB_TRAIN_EPI_STATS', '0') == '1'
        self.epi_stats_mink = int(os.environ.get('IFCB_TRAIN_EPI_STATS_MINK', '512'))
        n_stat = sum(2 * int(v.shape[0]) + 16 for k, v in sd.items() if v.dim() == 4)
        self.stat_sums = torch.zeros((n_stat,), dtype=torch.float64, device=self.device)
        self._stat_cursor = 0
        if arch == 'inception_v3':
            _build_inception_train(self, sd)
        elif arch in RESNET_CFG:
            _build_resnet_train(self, sd, arch)
        elif arch in PLAIN_ARCHS:
            _build_plain_train(self, sd, arch)
        elif arch == 'squeezenet':
            _build_squeezenet_train(self, sd)
        elif arch.startswith('densenet'):
            _build_densenet_train(self, sd)
        else:
            raise KeyError('model unknown!')
        self.n_params = self._cursor
        if share is not None:
            assert share.arch == arch and share.n_params == self.n_params, 'share: different model'
            self.grads, self.m, self.v = share.grads, share.m, share.v
        else:
            self.params = self.params[:self.n_params]
            self.grads = torch.zeros_like(self.params)
            self.m = torch.zeros_like(self.params)
            self.v = torch.zeros_like(self.params)
        self._plan_grad_borders()
        self._finalize(int(bucket_mb) << 20)
        if self.deterministic:
            self._det_ws = torch.zeros(self._det_need, dtype=torch.uint8, device=self.device)       # (torch allocations are 512-byte aligned)
            _lib.check(self.lib.ifcb_train_deterministic(self._det_ws.data_ptr(), self._det_ws.numel()), 'train_deterministic')
        self.repack()

    # ---- plumbing ---------------------------------------------------------------------------
    @property
    def step_count(self):
        return self._steps[0]

    @step_count.setter
    def step_count(self, v):
        self._steps[0] = int(v)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, name, *args):
        _lib.check(getattr(self.lib, name)(*args), name)

    def _param(self, name, init, kind, meta=None):
        n = init.numel()
        off = self._cursor
        assert off + n <= self.params.numel(), 'parameter arena overflow'
        if self._share is None:
            self.params[off:off + n].copy_(init.reshape(-1).float())
        self._cursor = ((off + n + 3) // 4) * 4                 # 16-byte aligned slices
        p = _Param(self, name, off, init.shape, kind, meta)
        self.plist.append(p)
        return p

    def _f32(self, n, fill=0.0):
        t = torch.full((n,), fill, dtype=torch.float32, device=self.device)
        self.keep.append(t)
        return t

    def alloc(self, H, W, Cc, pad=(0, 0)):
        return self.fp.alloc(H, W, Cc, pad)

    def border(self, H, W, Ci, Co, k, stride=(1, 1), pad=(0, 0)):
        """Zero border to allocate a tensor with when its consumer is conv(Ci -> Co, k, stride, pad) over H x W."""
        if not self.window:
            return (0, 0)
        return self.fp.border_for(H, W, Co, k, stride, pad, Ci=Ci)

    def grad_of(self, v):
        """Gradient view matching activation view ``v``.  The gradient tensor carries the zero border the
        data-gradient convs of v's producers want (``_grad_pad``, filled by ``_plan_grad_borders``)."""
        g = self._grad_t.get(id(v.t))
        pad = self._grad_pad.get(id(v.t), (0, 0))
        if g is None:
            g = torch.zeros((self.batch, v.H + 2 * pad[0], v.W + 2 * pad[1], v.t.shape[3]), dtype=self.tdtype, device=self.device)
            self._grad_t[id(v.t)] = g
            self.keep.append(v.t)
        return View(g, v.c0, v.c1, pad)

    def _plan_grad_borders(self):
        """For every conv output tensor: the border (k-1-pad) its stride-1 data-gradient conv reads as padding when
        the library will run that conv with the WINDOW scheme (max over the producers of a concat buffer)."""
        for rec in self.records:
            if rec['kind'] not in ('conv_bn', 'conv_act') or rec['stem'] is not None or rec['pool_after'] is not None or not self.window:
                continue
            if rec.get('group') is not None:
                continue
            if tuple(rec['stride']) != (1, 1):
                continue                                             # strided: the dilated copy carries the border
            kh, kw, pad, out = rec['kh'], rec['kw'], rec['pad'], rec['out']
            dpad = (kh - 1 - pad[0], kw - 1 - pad[1])
            want = self.bp.border_for(out.H, out.W, rec['Ci'], (kh, kw), pad=dpad, Ci=rec['Co'])
            cur = self._grad_pad.get(id(out.t), (0, 0))
            self._grad_pad[id(out.t)] = (max(cur[0], want[0]), max(cur[1], want[1]))

    # ---- graph construction (forward order) -------------------------------------------------------
    def conv_bn(self, x, sd, conv, bn, stride=(1, 1), pad=(0, 0), relu=True, residual=None, out=None, out_pad=(0, 0),
                eps=1e-5, stem=False, pool_after=None, conv_bias=False):
        """Conv2d(bias=False) -> BatchNorm2d(train) [-> + residual] [-> ReLU]; returns the activation view.

        ``pool_after=(k, stride, pad)``: the module computes avg_pool2d(x) -> 1x1 conv (Inception's
        branch_pool, inception.py:204,278,356); both are linear and commute, so the path evaluates
        conv -> avg_pool on the conv's 4-10x fewer output channels (forward and backward)."""
        w = sd[conv + '.weight'].float()
        Co, Ci, kh, kw = [int(s) for s in w.shape]
        stem_geom = None
        if stem:
            # Cin = 3: build the patch matrix once per step (ifcb_stem_im2col) and run the stem as a 1x1
            # convolution with Cin = K8 = round_up(kh*kw*3, 8) on the tensor-core forward / wgrad kernels
            assert Ci == 3 and stride[0] == stride[1] and pad[0] == pad[1]
            P = (self.R + 2 * pad[0] - kh) // stride[0] + 1
            K8 = ((kh * kw * 3 + 7) // 8) * 8
            x = self.alloc(P, P, K8)
            xd_ = _vd(x)
            stem_geom = dict(kh=kh, kw=kw, stride=stride, pad=pad)
            st, pd_, R_ = stride[0], pad[0], self.R
            tsc = tsh = None
            if self.transform_input:
                from .graph import transform_input_affine
                ts_, tb_ = transform_input_affine()
                tsc, tsh = (C.c_float * 3)(*ts_), (C.c_float * 3)(*tb_)
            self.fwd.append(lambda kh=kh, kw=kw: self._call('ifcb_stem_im2col', self.inp.data_ptr(), R_, R_, C.byref(xd_), self.batch, kh, kw,
                                                            st, pd_, tsc, tsh, self.cdtype, self._stream()))
            w1 = torch.zeros((Co, K8, 1, 1))
            w1[:, :kh * kw * 3, 0, 0] = w.permute(0, 2, 3, 1).reshape(Co, kh * kw * 3)      # k = (r*kw + s)*3 + c
            w, Ci, kh, kw, stride, pad = w1, K8, 1, 1, (1, 1), (0, 0)
        H, W = x.H, x.W
        P = (H + 2 * pad[0] - kh) // stride[0] + 1
        Q = (W + 2 * pad[1] - kw) // stride[1] + 1
        raw = None
        if pool_after is not None:
            assert (kh, kw) == (1, 1) and stride == (1, 1) and pad == (0, 0) and not stem
            raw = self.alloc(P, Q, Co)                               # conv output before the pool
            pk, ps, pp_ = pool_after
            P, Q = (P + 2 * pp_ - pk) // ps + 1, (Q + 2 * pp_ - pk) // ps + 1
        if out is None:
            out = self.alloc(P, Q, Co, out_pad)
        z = self.alloc(P, Q, Co)
        master = w.permute(0, 2, 3, 1).reshape(Co, kh * kw, Ci).contiguous()
        pw = self._param(conv + '.weight', master, 'stem' if stem else 'conv', dict(Ci=Ci, kh=kh, kw=kw, stem=stem_geom))
        # vgg*_bn: Conv2d(bias=True) in front of a BatchNorm.  The bias is added in the epilogue; its gradient is the per-channel sum
        # of the BatchNorm's dz, which is identically zero, so it is left at zero (torch's comes out as ~1e-9 rounding noise)
        pcb = self._param(conv + '.bias', sd[conv + '.bias'].float(), 'vec') if conv_bias else None
        pg = self._param(bn + '.weight', sd[bn + '.weight'], 'vec')
        pb = self._param(bn + '.bias', sd[bn + '.bias'], 'vec')
        if self._share is not None:
            rm, rv = self._share.buffers[bn + '.running_mean'], self._share.buffers[bn + '.running_var']
            self.buffers[bn + '.num_batches_tracked'] = self._share.buffers[bn + '.num_batches_tracked']
        else:
            rm = self._f32(Co); rm.copy_(sd[bn + '.running_mean'])
            rv = self._f32(Co); rv.copy_(sd[bn + '.running_var'])
            self.buffers[bn + '.num_batches_tracked'] = sd.get(bn + '.num_batches_tracked', torch.zeros((), dtype=torch.long)).clone()
        self.buffers[bn + '.running_mean'], self.buffers[bn + '.running_var'] = rm, rv
        mean, invstd = self._f32(Co), self._f32(Co)
        ones, zeros = torch.ones(Co), torch.zeros(Co)
        li = len(self.fp.layer_names)
        sums_ptr = None
        if self.epi_stats and raw is None and kh * kw * Ci >= self.epi_stats_mink:     # (a pooled branch normalises the POOLED tensor)
            sums_ptr = self.stat_sums.data_ptr() + 8 * self._stat_cursor
            self._stat_cursor += 2 * Co + (-(2 * Co) % 2)
        self.fp.conv(x, [dict(weight=w, scale=ones, shift=zeros, relu=False, out=z if raw is None else raw)], stride, pad, name=conv,
                     stats=sums_ptr, shift_ptr=pcb.wptr if pcb is not None else None)
        wf = self.fp.last_weight                                     # packed [Cout_pad, K_pad] 16-bit operand
        rawd = _vd(raw) if raw is not None else None
        zd, od = _vd(z), _vd(out)
        rd = _vd(residual) if residual is not None else None
        B, dt = self.batch, self.cdtype
        bn_key = bn + '.num_batches_tracked'

        def fwd():
            self.fp.run(B, li, li + 1)
            if rawd is not None:
                self._call('ifcb_avgpool_fwd', C.byref(rawd), C.byref(zd), B, pool_after[0], pool_after[1], pool_after[2], dt, self._stream())
            if sums_ptr is not None:
                self._call('ifcb_bn_apply_sums', C.byref(zd), C.byref(od), C.byref(rd) if rd is not None else None, B, dt, sums_ptr, eps, 0.1,
                           mean.data_ptr(), invstd.data_ptr(), rm.data_ptr(), rv.data_ptr(), pg.wptr, pb.wptr, 1 if relu else 0, self._stream())
                return
            self._call('ifcb_bn_stats', C.byref(zd), B, dt, eps, 0.1, self.acc.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                       rm.data_ptr(), rv.data_ptr(), self._stream())
            self._call('ifcb_bn_apply', C.byref(zd), C.byref(od), C.byref(rd) if rd is not None else None, B, dt, mean.data_ptr(),
                       invstd.data_ptr(), pg.wptr, pb.wptr, 1 if relu else 0, self._stream())
        self.fwd.append(fwd)
        self._nbt_keys.append(bn_key)
        self.records.append(dict(kind='conv_bn', x=x, z=z, out=out, residual=residual, relu=relu, stride=stride, pad=pad,
                                 pw=pw, pg=pg, pb=pb, mean=mean, invstd=invstd, wf=wf, Co=Co, Ci=Ci, kh=kh, kw=kw, stem=stem_geom,
                                 H=H, W=W, name=conv, raw=raw, pool_after=pool_after))
        return out

    def conv_bn_group(self, x, sd, specs, eps=1e-5, name='group'):
        """Sibling 1x1 conv + BN + ReLU units that read the SAME input (the branches of an Inception block): ONE forward GEMM with
        the filters concatenated along N, ONE weight-gradient launch and ONE data-gradient launch (K = all branches' channels)
        instead of one each per branch -- the block input is read once per pass instead of once per branch, and its gradient is
        written once instead of read-modify-written by every branch.  BatchNorm stays per branch (slices of the fused conv
        output / gradient tensors).  ``specs``: dicts(conv, bn, out=None, out_pad=(0, 0), pool_after=None); returns the
        activation views.  Records stay per branch (same fields as conv_bn), so the teacher-forced checker sees ordinary units."""
        B, dt = self.batch, self.cdtype
        ws = [sd[sp['conv'] + '.weight'].float() for sp in specs]
        assert all(tuple(w.shape[2:]) == (1, 1) and int(w.shape[1]) == x.C for w in ws)
        Cos, Ci, H, W = [int(w.shape[0]) for w in ws], x.C, x.H, x.W
        Ctot = sum(Cos)
        Zf = self.alloc(H, W, Ctot)                                  # fused conv output [.., sum Co]
        DZf = View(torch.zeros((B, H, W, Ctot), dtype=self.tdtype, device=self.device))     # its gradient
        self.keep.append(DZf.t)
        # the conv weights back to back in the arena (one [sum Co][1][Ci] block for the fused weight gradient), then the BN vectors
        pws = [self._param(sp['conv'] + '.weight', w.permute(0, 2, 3, 1).reshape(co, 1, Ci).contiguous(), 'conv',
                           dict(Ci=Ci, kh=1, kw=1, stem=None)) for sp, w, co in zip(specs, ws, Cos)]
        for a, b in zip(pws[:-1], pws[1:]):
            assert b.off == a.off + a.n, 'group weights must be contiguous'
        li = len(self.fp.layer_names)
        self.fp.conv(x, [dict(weight=torch.cat(ws, 0), scale=torch.ones(Ctot), shift=torch.zeros(Ctot), relu=False, out=Zf)], name=name)
        wf = self.fp.last_weight
        self.fwd.append(lambda: self.fp.run(B, li, li + 1))
        group = dict(x=x, Zf=Zf, DZf=DZf, pws=pws, Cos=Cos, Ci=Ci, wf=wf, H=H, W=W, n=len(specs), done=0, name=name)
        outs, c0 = [], 0
        for sp, co, pw in zip(specs, Cos, pws):
            bn, pool_after = sp['bn'], sp.get('pool_after')
            zs = Zf.slice(c0, c0 + co)
            raw, z = (zs, self.alloc(H, W, co)) if pool_after is not None else (None, zs)
            if pool_after is not None:
                assert tuple(pool_after) == (3, 1, 1)
            out = sp.get('out')
            if out is None:
                out = self.alloc(H, W, co, sp.get('out_pad', (0, 0)))
            pg = self._param(bn + '.weight', sd[bn + '.weight'], 'vec')
            pb = self._param(bn + '.bias', sd[bn + '.bias'], 'vec')
            if self._share is not None:
                rm, rv = self._share.buffers[bn + '.running_mean'], self._share.buffers[bn + '.running_var']
                self.buffers[bn + '.num_batches_tracked'] = self._share.buffers[bn + '.num_batches_tracked']
            else:
                rm = self._f32(co); rm.copy_(sd[bn + '.running_mean'])
                rv = self._f32(co); rv.copy_(sd[bn + '.running_var'])
                self.buffers[bn + '.num_batches_tracked'] = sd.get(bn + '.num_batches_tracked', torch.zeros((), dtype=torch.long)).clone()
            self.buffers[bn + '.running_mean'], self.buffers[bn + '.running_var'] = rm, rv
            mean, invstd = self._f32(co), self._f32(co)
            rawd = _vd(raw) if raw is not None else None
            zd, od = _vd(z), _vd(out)

            def fwd(rawd=rawd, zd=zd, od=od, mean=mean, invstd=invstd, rm=rm, rv=rv, pg=pg, pb=pb, pa=pool_after):
                if rawd is not None:
                    self._call('ifcb_avgpool_fwd', C.byref(rawd), C.byref(zd), B, pa[0], pa[1], pa[2], dt, self._stream())
                self._call('ifcb_bn_stats', C.byref(zd), B, dt, eps, 0.1, self.acc.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                           rm.data_ptr(), rv.data_ptr(), self._stream())
                self._call('ifcb_bn_apply', C.byref(zd), C.byref(od), None, B, dt, mean.data_ptr(), invstd.data_ptr(), pg.wptr, pb.wptr, 1,
                           self._stream())
            self.fwd.append(fwd)
            self._nbt_keys.append(bn + '.num_batches_tracked')
            self.records.append(dict(kind='conv_bn', x=x, z=z, out=out, residual=None, relu=True, stride=(1, 1), pad=(0, 0), pw=pw, pg=pg, pb=pb,
                                     mean=mean, invstd=invstd, wf=wf, Co=co, Ci=Ci, kh=1, kw=1, stem=None, H=H, W=W, name=sp['conv'],
                                     raw=raw, pool_after=pool_after, group=group, gslice=(c0, c0 + co)))
            outs.append(out)
            c0 += co
        return outs

    def conv_act(self, x, sd, conv, stride=(1, 1), pad=(0, 0), relu=True, out=None, out_pad=(0, 0), stem=False, bias=True, weight=None,
                 linear=False, co_pad=None):
        """Conv2d(+bias) [-> ReLU] WITHOUT BatchNorm (AlexNet / VGG / SqueezeNet convs and their Linear layers run as convs,
        DenseNet's raw convs).  The epilogue adds the bias and applies the ReLU; the backward pass masks the gradient with the
        stored output and sums it into the bias gradient (ifcb_bias_relu_backward).  ``weight``: override (a Linear weight
        viewed as a conv filter); ``co_pad``: pad the output channels with zero filters up to a multiple of 8."""
        w = (weight if weight is not None else sd[conv + '.weight']).float()
        b = sd[conv + '.bias'].float() if bias else None
        Co_real = int(w.shape[0])
        if co_pad is not None and co_pad > Co_real:
            w = torch.cat([w, torch.zeros((co_pad - Co_real,) + tuple(w.shape[1:]))], 0)
            if b is not None:
                b = torch.cat([b, torch.zeros(co_pad - Co_real)])
        Co, Ci, kh, kw = [int(v) for v in w.shape]
        stem_geom = None
        if stem:
            assert Ci == 3 and stride[0] == stride[1] and pad[0] == pad[1]
            P = (self.R + 2 * pad[0] - kh) // stride[0] + 1
            K8 = ((kh * kw * 3 + 7) // 8) * 8
            x = self.alloc(P, P, K8)
            xd_ = _vd(x)
            stem_geom = dict(kh=kh, kw=kw, stride=stride, pad=pad)
            st, pd_, R_ = stride[0], pad[0], self.R
            self.fwd.append(lambda kh=kh, kw=kw: self._call('ifcb_stem_im2col', self.inp.data_ptr(), R_, R_, C.byref(xd_), self.batch, kh, kw,
                                                            st, pd_, None, None, self.cdtype, self._stream()))
            w1 = torch.zeros((Co, K8, 1, 1))
            w1[:, :kh * kw * 3, 0, 0] = w.permute(0, 2, 3, 1).reshape(Co, kh * kw * 3)
            w, Ci, kh, kw, stride, pad = w1, K8, 1, 1, (1, 1), (0, 0)
        H, W = x.H, x.W
        P = (H + 2 * pad[0] - kh) // stride[0] + 1
        Q = (W + 2 * pad[1] - kw) // stride[1] + 1
        if out is None:
            out = self.alloc(P, Q, Co, out_pad)
        master = w.permute(0, 2, 3, 1).reshape(Co, kh * kw, Ci).contiguous()
        pw = self._param(conv + '.weight', master, 'stem' if stem else 'conv',
                         dict(Ci=Ci, kh=kh, kw=kw, stem=stem_geom, linear=linear, Co_real=Co_real))
        pb = self._param(conv + '.bias', b, 'vec', dict(Co_real=Co_real)) if b is not None else None
        li = len(self.fp.layer_names)
        self.fp.conv(x, [dict(weight=w, scale=torch.ones(Co), shift=torch.zeros(Co), relu=relu, out=out)], stride, pad, name=conv,
                     shift_ptr=pb.wptr if pb is not None else None)
        wf = self.fp.last_weight
        B = self.batch
        self.fwd.append(lambda: self.fp.run(B, li, li + 1))
        self.records.append(dict(kind='conv_act', x=x, out=out, relu=relu, stride=stride, pad=pad, pw=pw, pb=pb, wf=wf, Co=Co, Ci=Ci, kh=kh,
                                 kw=kw, stem=stem_geom, H=H, W=W, name=conv, pool_after=None, residual=None))
        return out

    def bn_act(self, x, sd, bn, relu=True, eps=1e-5, first=False):
        """BatchNorm2d(train) [-> ReLU] over an EXISTING tensor view (DenseNet's pre-activation norms over the growing
        concatenation).  ``first``: this is the first unit of the backward pass to write the input's gradient (it covers the
        whole tensor); every other norm ADDS its contribution (ifcb_bn_backward_accumulate)."""
        Cc = x.C
        out = self.alloc(x.H, x.W, Cc)
        pg = self._param(bn + '.weight', sd[bn + '.weight'], 'vec')
        pb = self._param(bn + '.bias', sd[bn + '.bias'], 'vec')
        if self._share is not None:
            rm, rv = self._share.buffers[bn + '.running_mean'], self._share.buffers[bn + '.running_var']
            self.buffers[bn + '.num_batches_tracked'] = self._share.buffers[bn + '.num_batches_tracked']
        else:
            rm = self._f32(Cc); rm.copy_(sd[bn + '.running_mean'])
            rv = self._f32(Cc); rv.copy_(sd[bn + '.running_var'])
            self.buffers[bn + '.num_batches_tracked'] = sd.get(bn + '.num_batches_tracked', torch.zeros((), dtype=torch.long)).clone()
        self.buffers[bn + '.running_mean'], self.buffers[bn + '.running_var'] = rm, rv
        mean, invstd = self._f32(Cc), self._f32(Cc)
        xd, od, B, dt = _vd(x), _vd(out), self.batch, self.cdtype

        def fwd():
            self._call('ifcb_bn_stats', C.byref(xd), B, dt, eps, 0.1, self.acc.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                       rm.data_ptr(), rv.data_ptr(), self._stream())
            self._call('ifcb_bn_apply', C.byref(xd), C.byref(od), None, B, dt, mean.data_ptr(), invstd.data_ptr(), pg.wptr, pb.wptr,
                       1 if relu else 0, self._stream())
        self.fwd.append(fwd)
        self._nbt_keys.append(bn + '.num_batches_tracked')
        self.records.append(dict(kind='bn_act', x=x, out=out, relu=relu, pg=pg, pb=pb, mean=mean, invstd=invstd, first=first, name=bn))
        return out

    def drop(self, x, p, tag):
        """nn.Dropout(p) in train mode on an activation tensor (classifier stacks of AlexNet / VGG / SqueezeNet): the mask times
        1/(1-p) comes from the counter-based generator (ifcb_dropout_scale) once per step; identity when dropout is switched off."""
        if not self.dropout or p <= 0:
            return x
        out = self.alloc(x.H, x.W, x.C)
        n = self.batch * x.H * x.W * x.C
        scale = self._f32(n, 1.0)
        xd, od, B, dt = _vd(x), _vd(out), self.batch, self.cdtype
        salt = len(self.pre) + 7

        def pre():
            seed = (self.seed * 1000003 + self.step_count * 7919 + salt) & ((1 << 63) - 1)
            self._call('ifcb_dropout_scale', scale.data_ptr(), n, p, seed, self._stream())
        self.pre.append(pre)
        self.fwd.append(lambda: self._call('ifcb_scale_elems', C.byref(xd), C.byref(od), scale.data_ptr(), B, dt, self._stream()))
        self.records.append(dict(kind='drop', x=x, out=out, scale=scale, name=tag))
        return out

    def maxpool(self, x, k, stride, pad, out=None, out_pad=(0, 0), ceil_mode=False):
        P = (x.H + 2 * pad - k) // stride + 1
        Q = (x.W + 2 * pad - k) // stride + 1
        if ceil_mode:                                  # torch: one more window if it still starts inside the input
            ce = lambda n: (lambda o: o - 1 if (o - 1) * stride >= n + pad else o)((n + 2 * pad - k + stride - 1) // stride + 1)
            P, Q = ce(x.H), ce(x.W)
        if out is None:
            out = self.alloc(P, Q, x.C, out_pad)
        idx = torch.zeros((self.batch, P, Q, x.C), dtype=torch.uint8, device=self.device)
        self.keep.append(idx)
        xd, od, B, dt = _vd(x), _vd(out), self.batch, self.cdtype
        self.fwd.append(lambda: self._call('ifcb_maxpool_fwd_train', C.byref(xd), C.byref(od), idx.data_ptr(), B, k, stride, pad, dt,
                                           self._stream()))
        self.records.append(dict(kind='maxpool', x=x, out=out, idx=idx, k=k, stride=stride, pad=pad))
        return out

    def avgpool(self, x, k, stride, pad, out=None):
        P = (x.H + 2 * pad - k) // stride + 1
        Q = (x.W + 2 * pad - k) // stride + 1
        if out is None:
            out = self.alloc(P, Q, x.C)
        xd, od, B, dt = _vd(x), _vd(out), self.batch, self.cdtype
        self.fwd.append(lambda: self._call('ifcb_avgpool_fwd', C.byref(xd), C.byref(od), B, k, stride, pad, dt, self._stream()))
        self.records.append(dict(kind='avgpool', x=x, out=out, k=k, stride=stride, pad=pad))
        return out

    def head(self, x, sd, fc, loss_weight=1.0, dropout_p=0.0, which='main', identity_classes=None):
        """adaptive_avg_pool2d(1) -> dropout -> Linear -> loss_weight * CrossEntropyLoss.  ``identity_classes`` = n: there is no
        Linear (SqueezeNet: the pooled outputs of the last conv ARE the logits) -- a fixed identity matrix stands in, outside the
        parameter arena, and its gradient is discarded."""
        if identity_classes is not None:
            n_classes, Cc = int(identity_classes), x.C
            eye = torch.zeros((n_classes, Cc)); eye[:, :n_classes] = torch.eye(n_classes)
            wt_, gt_ = self._f32(n_classes * Cc), self._f32(n_classes * Cc + n_classes)
            wt_.copy_(eye.reshape(-1))
            bz_ = self._f32(n_classes)

            class _Fixed(object):                      # quacks like a _Param for the head kernels
                def __init__(s_, w, g): s_.wptr, s_.gptr, s_.off = w, g, None
            pw, pb = _Fixed(wt_.data_ptr(), gt_.data_ptr()), _Fixed(bz_.data_ptr(), gt_.data_ptr() + 4 * n_classes * Cc)
        else:
            Wt, bias = sd[fc + '.weight'].float(), sd[fc + '.bias'].float()
            n_classes, Cc = int(Wt.shape[0]), int(Wt.shape[1])
            assert Cc == x.C
            pw = self._param(fc + '.weight', Wt, 'fc')
            pb = self._param(fc + '.bias', bias, 'vec')
        B, dt = self.batch, self.cdtype
        pooled = self._f32(B * Cc)
        logits = self._f32(B * n_classes)
        dlogits = self._f32(B * n_classes)
        drop = self._f32(B * Cc, 1.0) if dropout_p > 0 else None
        xd = _vd(x)
        rec = dict(kind='head', x=x, pw=pw, pb=pb, pooled=pooled, dlogits=dlogits, logits=logits, drop=drop, n_classes=n_classes,
                   dropout_p=dropout_p, which=which)

        def pre():
            if drop is not None and self.dropout and not rec.get('fixed_mask'):
                seed = (self.seed * 1000003 + self.step_count * 7919 + (1 if which == 'aux' else 0)) & ((1 << 63) - 1)
                self._call('ifcb_dropout_scale', drop.data_ptr(), B * Cc, dropout_p, seed, self._stream())
        self.pre.append(pre)

        def fwd():
            use_drop = drop is not None and (self.dropout or rec.get('fixed_mask'))
            self._call('ifcb_head_train_fwd', C.byref(xd), B, dt, drop.data_ptr() if use_drop else None, pw.wptr, pb.wptr,
                       self.labels.data_ptr(), n_classes, loss_weight, pooled.data_ptr(), logits.data_ptr(), dlogits.data_ptr(),
                       self.loss.data_ptr(), self._stream())
        self.fwd.append(fwd)
        self.records.append(rec)
        if which == 'main':
            self.logits = logits.view(B, n_classes)
            self.n_classes = n_classes
            self.main_head = rec
        return rec

    # ---- backward construction (reverse order) ------------------------------------------------------
    def _finalize(self, bucket_bytes):
        B, dt = self.batch, self.cdtype
        written = set()

        def claim(v):
            """True if a gradient was already written into v's tensor slice (=> accumulate)."""
            key = (id(v.t), v.c0, v.c1)
            for (t, a, b) in written:
                if t == key[0] and (a, b) != (key[1], key[2]) and a < key[2] and key[1] < b:
                    raise RuntimeError('partially overlapping gradient consumers are not supported')
            acc = key in written
            written.add(key)
            return acc

        unit_lo = []                    # per backward closure: arena offset its unit's parameters start at (None: none)
        for rec in reversed(self.records):
            kind = rec['kind']
            if kind == 'head':
                x = rec['x']
                dx = self.grad_of(x)
                acc = claim(x)
                dxd = _vd(dx)
                use = rec

                def bwd(rec=rec, dxd=dxd, acc=acc):
                    use_drop = rec['drop'] is not None and (self.dropout or rec.get('fixed_mask'))
                    self._call('ifcb_head_bwd', C.byref(dxd), 1 if acc else 0, B, dt, rec['drop'].data_ptr() if use_drop else None,
                               rec['pw'].wptr, rec['pooled'].data_ptr(), rec['dlogits'].data_ptr(), rec['n_classes'],
                               rec['pw'].gptr, rec['pb'].gptr, self._stream())
                self.bwd.append(bwd)
                lo = rec['pw'].off
            elif kind in ('maxpool', 'avgpool'):
                x, out = rec['x'], rec['out']
                dy, dx = self.grad_of(out), self.grad_of(x)
                acc = claim(x)
                dyd, dxd = _vd(dy), _vd(dx)
                k, s, p = rec['k'], rec['stride'], rec['pad']
                if kind == 'maxpool':
                    idx = rec['idx']
                    self.bwd.append(lambda dyd=dyd, dxd=dxd, idx=idx, acc=acc, k=k, s=s, p=p: self._call(
                        'ifcb_maxpool_bwd', C.byref(dyd), idx.data_ptr(), C.byref(dxd), 1 if acc else 0, B, k, s, p, dt, self._stream()))
                else:
                    self.bwd.append(lambda dyd=dyd, dxd=dxd, acc=acc, k=k, s=s, p=p: self._call(
                        'ifcb_avgpool_bwd', C.byref(dyd), C.byref(dxd), 1 if acc else 0, B, k, s, p, dt, self._stream()))
                lo = None
            elif kind == 'conv_act':
                out, pb, relu = rec['out'], rec['pb'], rec['relu']
                dy = self.grad_of(out)
                rec['dy'] = rec['dz'] = dy
                if pb is not None or relu:
                    dyd, od = _vd(dy), _vd(out)
                    if pb is None:
                        pb = rec['pb_dummy'] = self._f32(rec['Co'])
                        gptr = pb.data_ptr()
                    else:
                        gptr = pb.gptr
                    self.bwd.append(lambda dyd=dyd, od=od, relu=relu, gptr=gptr: self._call(
                        'ifcb_bias_relu_backward', C.byref(dyd), C.byref(od) if relu else None, C.byref(dyd), 1 if relu else 0, B, dt,
                        self.acc.data_ptr(), gptr, self._stream()))
                lo = self._conv_grads(rec, dy, claim)
            elif kind == 'bn_act':
                x, out, relu = rec['x'], rec['out'], rec['relu']
                dy, dx = self.grad_of(out), self.grad_of(x)
                dyd, xd, dxd = _vd(dy), _vd(x), _vd(dx)
                pg, pb_, mean, invstd = rec['pg'], rec['pb'], rec['mean'], rec['invstd']
                if rec['first']:          # covers the whole tensor and comes first in the backward pass: plain write
                    self.bwd.append(lambda dyd=dyd, xd=xd, dxd=dxd, relu=relu, pg=pg, pb_=pb_, mean=mean, invstd=invstd: self._call(
                        'ifcb_bn_backward', C.byref(dyd), None, C.byref(xd), C.byref(dxd), None, 0, 1 if relu else 0, B, dt, mean.data_ptr(),
                        invstd.data_ptr(), pg.wptr, pb_.wptr, self.acc.data_ptr(), pg.gptr, pb_.gptr, self._stream()))
                else:
                    self.bwd.append(lambda dyd=dyd, xd=xd, dxd=dxd, relu=relu, pg=pg, pb_=pb_, mean=mean, invstd=invstd: self._call(
                        'ifcb_bn_backward_accumulate', C.byref(dyd), C.byref(xd), C.byref(dxd), 1 if relu else 0, B, dt, mean.data_ptr(),
                        invstd.data_ptr(), pg.wptr, pb_.wptr, self.acc.data_ptr(), pg.gptr, pb_.gptr, self._stream()))
                lo = pg.off
            elif kind == 'drop':
                x, out, scale = rec['x'], rec['out'], rec['scale']
                dy, dx = self.grad_of(out), self.grad_of(x)
                assert not claim(x), 'dropout input with several consumers'
                dyd, dxd = _vd(dy), _vd(dx)
                self.bwd.append(lambda dyd=dyd, dxd=dxd, scale=scale: self._call('ifcb_scale_elems', C.byref(dyd), C.byref(dxd), scale.data_ptr(),
                                                                                 B, dt, self._stream()))
                lo = None
            else:
                lo = self._finalize_conv(rec, claim)
            unit_lo.extend([None] * (len(self.bwd) - len(unit_lo) - 1) + [lo])
        # contiguous arena ranges to all-reduce as soon as the backward pass has finished them
        self.bucket_marks = plan_buckets(unit_lo, self.n_params, bucket_bytes // 4)

    def _finalize_group_member(self, rec, claim):
        """BatchNorm backward of one branch of a conv_bn_group into its slice of the fused gradient tensor; after the last branch
        (the first in forward order) the fused weight-gradient and data-gradient launches."""
        B, dt = self.batch, self.cdtype
        g, (c0, c1) = rec['group'], rec['gslice']
        z, out, Co = rec['z'], rec['out'], rec['Co']
        dy = self.grad_of(out)
        target = g['DZf'].slice(c0, c1)
        pa = rec['pool_after']
        if pa is not None:                                           # BN's dz is w.r.t. the POOLED tensor: own buffer, then avg-pool backward
            dz = View(torch.zeros((B, z.H, z.W, Co), dtype=self.tdtype, device=self.device))
            self.keep.append(dz.t)
        else:
            dz = target
        rec['dy'], rec['dz'] = dy, dz
        dyd, zd, dzd, td = _vd(dy), _vd(z), _vd(dz), _vd(target)
        pg, pb, mean, invstd = rec['pg'], rec['pb'], rec['mean'], rec['invstd']

        def bn_bwd():
            self._call('ifcb_bn_backward', C.byref(dyd), None, C.byref(zd), C.byref(dzd), None, 0, 1, B, dt, mean.data_ptr(), invstd.data_ptr(),
                       pg.wptr, pb.wptr, self.acc.data_ptr(), pg.gptr, pb.gptr, self._stream())
            if pa is not None:
                self._call('ifcb_avgpool_bwd', C.byref(dzd), C.byref(td), 0, B, pa[0], pa[1], pa[2], dt, self._stream())
        self.bwd.append(bn_bwd)
        if pa is not None:
            rec['dr'] = target
        g['done'] += 1
        if g['done'] < g['n']:
            return pg.off
        # ---- all branches have written their slice of DZf: fused weight gradient + data gradient ----
        x, Ci, Ctot, pws = g['x'], g['Ci'], sum(g['Cos']), g['pws']
        DZf = g['DZf']
        wd = WgradDesc()
        wd.d_in, wd.in_ld, wd.Cin = x.ptr, x.ld, x.C
        wd.batch, wd.H, wd.W = B, g['H'], g['W']
        wd.in_pad_h, wd.in_pad_w = x.pad
        wd.kh, wd.kw, wd.stride_h, wd.stride_w, wd.pad_h, wd.pad_w = 1, 1, 1, 1, 0, 0
        wd.d_dout, wd.dout_ld, wd.Cout = DZf.ptr, DZf.ld, Ctot
        wd.dout_pad_h, wd.dout_pad_w = DZf.pad
        wd.dtype = dt
        if self.deterministic:
            self._det_need = max(self._det_need, int(self.lib.ifcb_conv_wgrad_workspace_bytes(C.byref(wd))))
        gptr0 = pws[0].gptr

        def wgrad():
            wd.d_dweight = gptr0
            self._call('ifcb_conv_wgrad', C.byref(wd), self._stream())
        self.bwd.append(wgrad)
        dx = self.grad_of(x)
        acc = claim(x)
        dg = build_dgrad(self.bp, DZf, dx, Ctot, Ci, 1, 1, (1, 1), (0, 0), acc, name='dgrad.' + g['name'])
        self.bwd.extend(dg['run'])
        geo_f = _lib.conv_geometry(Ci, Ctot, 1, 1)
        wf, wdg = g['wf'], dg['weight']
        assert wf.shape[1] == geo_f['Cin_pad']
        co0 = 0
        for pw, co in zip(pws, g['Cos']):
            self.repacks.append((pw, co, 1, Ci, wf.data_ptr() + 2 * co0 * geo_f['Cin_pad'], geo_f['Cin_pad'],
                                 wdg.data_ptr() + 2 * co0, dg['Cin_pad']))
            co0 += co
        self.keep.extend([wf, wdg])
        return pws[0].off

    def _finalize_conv(self, rec, claim):
        if rec.get('group') is not None:
            return self._finalize_group_member(rec, claim)
        B, dt = self.batch, self.cdtype
        x, z, out, residual = rec['x'], rec['z'], rec['out'], rec['residual']
        Co, Ci, kh, kw = rec['Co'], rec['Ci'], rec['kh'], rec['kw']
        stride, pad = rec['stride'], rec['pad']
        dy = self.grad_of(out)
        dyd, zd, od = _vd(dy), _vd(z), _vd(out)
        if self.keep_dy:
            gp = dy.pad
            dz = View(torch.zeros((B, out.H + 2 * gp[0], out.W + 2 * gp[1], Co), dtype=self.tdtype, device=self.device), pad=gp)
            self.keep.append(dz.t)
        else:
            dz = dy                                                  # in place
        dzd = _vd(dz)
        rec['dy'], rec['dz'] = dy, dz
        dres_d, res_acc = None, False
        if residual is not None:
            dres = self.grad_of(residual)
            res_acc = claim(residual)
            dres_d = _vd(dres)
        pw, pg, pb, mean, invstd = rec['pw'], rec['pg'], rec['pb'], rec['mean'], rec['invstd']
        relu = rec['relu']

        def bn_bwd():
            self._call('ifcb_bn_backward', C.byref(dyd), C.byref(od) if (relu and dres_d is not None) else None, C.byref(zd),
                       C.byref(dzd), C.byref(dres_d) if dres_d is not None else None, 1 if res_acc else 0, 1 if relu else 0, B, dt,
                       mean.data_ptr(), invstd.data_ptr(), pg.wptr, pb.wptr, self.acc.data_ptr(), pg.gptr, pb.gptr, self._stream())
        self.bwd.append(bn_bwd)
        if rec['pool_after'] is not None:                            # d(conv output) = avg_pool backward of dz
            raw = rec['raw']
            dr = View(torch.zeros((B, raw.H, raw.W, Co), dtype=self.tdtype, device=self.device))
            self.keep.append(dr.t)
            drd, pa = _vd(dr), rec['pool_after']
            self.bwd.append(lambda: self._call('ifcb_avgpool_bwd', C.byref(dzd), C.byref(drd), 0, B, pa[0], pa[1], pa[2], dt, self._stream()))
            rec['dr'] = dr
            dz = dr
        return self._conv_grads(rec, dz, claim)

    def _conv_grads(self, rec, dz, claim):
        """Weight gradient and data gradient of a conv unit from the gradient ``dz`` of its (pre-activation) output."""
        B, dt = self.batch, self.cdtype
        x, pw = rec['x'], rec['pw']
        Co, Ci, kh, kw = rec['Co'], rec['Ci'], rec['kh'], rec['kw']
        stride, pad = rec['stride'], rec['pad']
        # weight gradient: dW[co, tap, ci] += sum dz * x
        wd = WgradDesc()
        xin = x
        wd.d_in, wd.in_ld, wd.Cin = xin.ptr, xin.ld, xin.C
        wd.batch, wd.H, wd.W = B, rec['H'], rec['W']
        wd.in_pad_h, wd.in_pad_w = xin.pad
        wd.kh, wd.kw, wd.stride_h, wd.stride_w, wd.pad_h, wd.pad_w = kh, kw, stride[0], stride[1], pad[0], pad[1]
        wd.d_dout, wd.dout_ld, wd.Cout = dz.ptr, dz.ld, Co
        wd.dout_pad_h, wd.dout_pad_w = dz.pad
        wd.dtype = dt

        if self.deterministic:
            self._det_need = max(self._det_need, int(self.lib.ifcb_conv_wgrad_workspace_bytes(C.byref(wd))))

        def wgrad():
            wd.d_dweight = pw.gptr
            self._call('ifcb_conv_wgrad', C.byref(wd), self._stream())
        self.bwd.append(wgrad)
        geo_f = _lib.conv_geometry(Ci, Co, kh, kw)
        wf = rec['wf']
        assert wf.shape[1] == kh * kw * geo_f['Cin_pad'], wf.shape
        if rec['stem'] is not None:                                  # network input: no data gradient
            self.repacks.append((pw, Co, kh * kw, Ci, wf, geo_f['Cin_pad'], None, 0))
            return pw.off
        # data gradient: stride-1 conv of (dilated) dz with the flipped, transposed filter
        dx = self.grad_of(x)
        acc = claim(x)
        dg = build_dgrad(self.bp, dz, dx, Co, Ci, kh, kw, stride, pad, acc, name='dgrad.' + rec['name'])
        self.bwd.extend(dg['run'])
        wdg = dg['weight']
        self.repacks.append((pw, Co, kh * kw, Ci, wf, geo_f['Cin_pad'], wdg, dg['Cin_pad']))
        return pw.off

    # ---- execution ---------------------------------------------------------------------------------
    def repack(self):
        """Refresh every 16-bit tensor-core operand (forward + data-gradient) from the fp32 arena: one launch over a
        device-resident table of (master slice, operands, geometry) records."""
        if self._repack_table is None:
            import numpy as np
            items = (_lib.RepackItem * len(self.repacks))()
            for it, (pw, Co, taps, Ci, wf, cin_pad, wdg, cout_padk) in zip(items, self.repacks):
                ptr = lambda t: t if (t is None or isinstance(t, int)) else t.data_ptr()
                it.d_master, it.d_wfwd, it.d_wdgrad = pw.wptr, ptr(wf), ptr(wdg)
                it.Cout, it.taps, it.Cin, it.Cin_pad, it.Cout_padk = Co, taps, Ci, cin_pad, cout_padk
            raw = np.frombuffer(bytes(items), dtype=np.uint8).copy()
            self._repack_table = torch.from_numpy(raw).to(self.device)
        self._call('ifcb_conv_repack_batch', self._repack_table.data_ptr(), len(self.repacks), self.cdtype, self._stream())

    def _count_batch(self):
        for k in self._nbt_keys:
            self.buffers[k] += 1

    def forward(self):
        for f in self.pre:
            f()
        self._forward_kernels()
        self._count_batch()

    def _forward_kernels(self):
        self._call('ifcb_memset_zero', self.loss.data_ptr(), 8, self._stream())
        if self._stat_cursor:
            self._call('ifcb_memset_zero', self.stat_sums.data_ptr(), 8 * self._stat_cursor, self._stream())
        for f in self.fwd:
            f()

    def enable_cuda_graph(self):
        """Captures the step's kernels into CUDA graphs: forward + backward split into one graph per gradient bucket
        (the NCCL all-reduce of a finished bucket is launched eagerly between replays, so a data-parallel step
        keeps its overlap), plus one graph for the operand repack.  Kernel arguments that change per step stay
        outside: the dropout seed (``pre`` closures) and Adam's bias corrections.  Every pointer is fixed at
        construction, so replays are valid for the life of the object."""
        saved = {k: v.clone() for k, v in self.buffers.items() if v.is_cuda}     # warm-up passes must not move the running statistics
        for _ in range(2):                                        # warm-up: function attributes, lazy module loads
            self._forward_kernels()
            self.backward()
            self.repack()
        for k, v in saved.items():
            self.buffers[k].copy_(v)
        torch.cuda.synchronize(self.device)
        marks = {}
        for m in self.bucket_marks:
            marks.setdefault(m[0], []).append((m[1], m[2]))
        cuts = sorted(marks)
        segs, start = [], 0
        for ci, cut in enumerate(cuts + ([len(self.bwd) - 1] if (not cuts or cuts[-1] != len(self.bwd) - 1) else [])):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if start == 0:
                    self._forward_kernels()
                    self._call('ifcb_memset_zero', self.grads.data_ptr(), 4 * self.n_params, self._stream())
                for b in self.bwd[start:cut + 1]:
                    b()
            segs.append((g, marks.get(cut, [])))
            start = cut + 1
        g_rp = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_rp):
            self.repack()
        self._graph = (segs, g_rp)

    def backward(self, allreduce=None):
        """Runs the backward pass; ``allreduce(lo, hi)`` is called as soon as gradient range [lo, hi) is final."""
        self._call('ifcb_memset_zero', self.grads.data_ptr(), 4 * self.n_params, self._stream())
        marks = {}
        for m in self.bucket_marks:
            marks.setdefault(m[0], []).append((m[1], m[2]))
        for i, b in enumerate(self.bwd):
            b()
            if allreduce is not None:
                for lo, hi in marks.get(i, ()):
                    allreduce(lo, hi)

    def forward_backward(self, x=None, labels=None, allreduce=None):
        if x is not None:
            self.inp.copy_(x)
        if labels is not None:
            self.labels.copy_(labels)
        self.forward()
        self.backward(allreduce)
        return self.loss[0]

    def adam(self, grad_scale=1.0):
        self.step_count += 1
        self._call('ifcb_adam_step', self.params.data_ptr(), self.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                   self.n_params, self.lr, self.betas[0], self.betas[1], self.eps, self.step_count, float(grad_scale), self._stream())
        self.repack()

    def step(self, x=None, labels=None):
        """training_step + backward + (gradient mean over ranks) + Adam.  Returns the loss (device scalar)."""
        if self._reducer is None:
            self._reducer = GradReducer(self.grads)
        if self._graph is not None:
            if x is not None:
                self.inp.copy_(x)
            if labels is not None:
                self.labels.copy_(labels)
            for f in self.pre:
                f()
            for g, ranges in self._graph[0]:
                g.replay()
                for lo, hi in ranges:
                    self._reducer(lo, hi)
            self._count_batch()
            self.step_count += 1
            self._call('ifcb_adam_step', self.params.data_ptr(), self.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                       self.n_params, self.lr, self.betas[0], self.betas[1], self.eps, self.step_count, float(self._reducer.wait()),
                       self._stream())
            self._graph[1].replay()
            return self.loss[0]
        loss = self.forward_backward(x, labels, self._reducer if self._reducer.world > 1 else None)
        self.adam(self._reducer.wait())
        return loss

    # ---- parameter import / export (torchvision layout) -----------------------------------------------
    def _export(self, arena):
        out = {}
        for p in self.plist:
            t = arena[p.off:p.off + p.n].view(p.shape)
            if p.kind == 'conv':
                Ci, kh, kw = p.meta['Ci'], p.meta['kh'], p.meta['kw']
                t = t[:, :, :Ci].reshape(p.shape[0], kh, kw, Ci).permute(0, 3, 1, 2)
                t = t[:p.meta.get('Co_real', p.shape[0])]
                if p.meta.get('linear'):                         # a Linear layer run as a conv over its whole input map
                    t = t.reshape(t.shape[0], -1)
            elif p.kind == 'vec' and 'Co_real' in p.meta:
                t = t[:p.meta['Co_real']]
            elif p.kind == 'stem':                               # [Co, 1, K8], k = (r*kw + s)*3 + c
                g = p.meta['stem']
                t = t[:, 0, :g['kh'] * g['kw'] * 3].reshape(p.shape[0], g['kh'], g['kw'], 3).permute(0, 3, 1, 2)
                t = t[:p.meta.get('Co_real', p.shape[0])]
            out[p.name] = t.detach().clone().contiguous()
        return out

    def state_dict(self):
        out = self._export(self.params)
        for k, v in self.buffers.items():
            out[k] = v.detach().clone()
        return {k: out[k].cpu() for k in self.sd_keys if k in out}

    def grad_dict(self):
        return self._export(self.grads)


# =====================================================================================================
# graphs (train mode; same module order / names as torchvision so that parameters line up)
# =====================================================================================================
def _build_resnet_train(tn, sd, arch):
    kind, layers = RESNET_CFG[arch]
    a = tn.conv_bn(None, sd, 'conv1', 'bn1', (2, 2), (3, 3), stem=True)
    H = (a.H + 2 - 3) // 2 + 1
    blocks = [(li, bi) for li, nb in enumerate(layers) for bi in range(nb)]
    # a tensor carries the zero border its 3x3 stride-1 consumer wants (WINDOW scheme); basic blocks: block inputs
    x = tn.maxpool(a, 3, 2, 1, out_pad=tn.border(H, H, 64, 64, (3, 3), pad=(1, 1)) if kind == 'basic' else (0, 0))
    for idx, (li, bi) in enumerate(blocks):
        pre = 'layer%d.%d' % (li + 1, bi)
        s = 2 if (li > 0 and bi == 0) else 1
        width = 64 << li
        Ho = (x.H + 2 - 3) // s + 1
        identity = x
        if (pre + '.downsample.0.weight') in sd:
            identity = tn.conv_bn(x, sd, pre + '.downsample.0', pre + '.downsample.1', (s, s), (0, 0), relu=False)
        if kind == 'basic':
            opad = (0, 0)
            if idx + 1 < len(blocks):
                nli, nbi = blocks[idx + 1]
                ns = 2 if (nli > 0 and nbi == 0) else 1
                opad = tn.border(Ho, Ho, width, 64 << nli, (3, 3), (ns, ns), (1, 1))
            t = tn.conv_bn(x, sd, pre + '.conv1', pre + '.bn1', (s, s), (1, 1), out_pad=tn.border(Ho, Ho, width, width, (3, 3), pad=(1, 1)))
            x = tn.conv_bn(t, sd, pre + '.conv2', pre + '.bn2', (1, 1), (1, 1), residual=identity, out_pad=opad)
        else:
            t = tn.conv_bn(x, sd, pre + '.conv1', pre + '.bn1', out_pad=tn.border(x.H, x.W, width, width, (3, 3), (s, s), (1, 1)))
            t = tn.conv_bn(t, sd, pre + '.conv2', pre + '.bn2', (s, s), (1, 1))
            x = tn.conv_bn(t, sd, pre + '.conv3', pre + '.bn3', residual=identity)
    tn.head(x, sd, 'fc')


def _build_inception_train(tn, sd):
    eps = 1e-3

    def cb(x, prefix, stride=(1, 1), pad=(0, 0), out=None, stem=False, pool_after=None, nxt=None):
        """``nxt`` = (Cout, k, pad) of the stride-1 conv that consumes the output: the output is allocated with the
        zero border that conv's WINDOW scheme reads as padding."""
        out_pad = (0, 0)
        if nxt is not None and out is None:
            w = sd[prefix + '.conv.weight']
            kh, kw = int(w.shape[2]), int(w.shape[3])
            Hin = tn.R if stem else x.H
            Ho = (Hin + 2 * pad[0] - kh) // stride[0] + 1
            Wo = (Hin + 2 * pad[1] - kw) // stride[1] + 1
            out_pad = tn.border(Ho, Wo, int(w.shape[0]), nxt[0], nxt[1], pad=nxt[2])
        return tn.conv_bn(x, sd, prefix + '.conv', prefix + '.bn', stride, pad, out=out, eps=eps, stem=stem, pool_after=pool_after,
                          out_pad=out_pad)

    a = cb(None, 'Conv2d_1a_3x3', (2, 2), stem=True)
    a = cb(a, 'Conv2d_2a_3x3', nxt=(64, (3, 3), (1, 1)))
    a = cb(a, 'Conv2d_2b_3x3', pad=(1, 1))
    a = tn.maxpool(a, 3, 2, 0)
    a = cb(a, 'Conv2d_3b_1x1')
    a = cb(a, 'Conv2d_4a_3x3')
    x = tn.maxpool(a, 3, 2, 0)
    fuse = os.environ.get('IFCB_TRAIN_FUSE_SIBLINGS', '1') != '0'        # sibling 1x1 convs of a block as one GEMM (A/B switch)

    def spec(prefix, out=None, nxt=None, pool_after=None, Hh=None):
        sp = dict(conv=prefix + '.conv', bn=prefix + '.bn', out=out, pool_after=pool_after)
        if nxt is not None and out is None:
            sp['out_pad'] = tn.border(Hh, Hh, int(sd[prefix + '.conv.weight'].shape[0]), nxt[0], nxt[1], pad=nxt[2])
        return sp

    for blk, pf in (('Mixed_5b', 32), ('Mixed_5c', 64), ('Mixed_5d', 64)):
        H = x.H
        out = tn.alloc(H, H, 224 + pf)
        if fuse:
            _, t5, t3, _ = tn.conv_bn_group(x, sd, [spec(blk + '.branch1x1', out=out.slice(0, 64)),
                                                    spec(blk + '.branch5x5_1', nxt=(64, (5, 5), (2, 2)), Hh=H),
                                                    spec(blk + '.branch3x3dbl_1', nxt=(96, (3, 3), (1, 1)), Hh=H),
                                                    spec(blk + '.branch_pool', out=out.slice(224, 224 + pf), pool_after=(3, 1, 1))],
                                            eps=eps, name=blk + '.1x1s')
            cb(t5, blk + '.branch5x5_2', pad=(2, 2), out=out.slice(64, 128))
            t = cb(t3, blk + '.branch3x3dbl_2', pad=(1, 1), nxt=(96, (3, 3), (1, 1)))
            cb(t, blk + '.branch3x3dbl_3', pad=(1, 1), out=out.slice(128, 224))
            x = out
            continue
        cb(x, blk + '.branch1x1', out=out.slice(0, 64))
        t = cb(x, blk + '.branch5x5_1', nxt=(64, (5, 5), (2, 2)))
        cb(t, blk + '.branch5x5_2', pad=(2, 2), out=out.slice(64, 128))
        t = cb(x, blk + '.branch3x3dbl_1', nxt=(96, (3, 3), (1, 1)))
        t = cb(t, blk + '.branch3x3dbl_2', pad=(1, 1), nxt=(96, (3, 3), (1, 1)))
        cb(t, blk + '.branch3x3dbl_3', pad=(1, 1), out=out.slice(128, 224))
        cb(x, blk + '.branch_pool', out=out.slice(224, 224 + pf), pool_after=(3, 1, 1))
        x = out
    blk = 'Mixed_6a'
    H2 = (x.H - 3) // 2 + 1
    out = tn.alloc(H2, H2, 768)
    cb(x, blk + '.branch3x3', (2, 2), out=out.slice(0, 384))
    t = cb(x, blk + '.branch3x3dbl_1', nxt=(96, (3, 3), (1, 1)))
    t = cb(t, blk + '.branch3x3dbl_2', pad=(1, 1))
    cb(t, blk + '.branch3x3dbl_3', (2, 2), out=out.slice(384, 480))
    tn.maxpool(x, 3, 2, 0, out=out.slice(480, 768))
    x = out
    for blk, c7 in (('Mixed_6b', 128), ('Mixed_6c', 160), ('Mixed_6d', 160), ('Mixed_6e', 192)):
        H = x.H
        out = tn.alloc(H, H, 768)
        w17, w71 = (1, 7), (7, 1)
        if fuse:
            _, t7, td, _ = tn.conv_bn_group(x, sd, [spec(blk + '.branch1x1', out=out.slice(0, 192)),
                                                    spec(blk + '.branch7x7_1', nxt=(c7, w17, (0, 3)), Hh=H),
                                                    spec(blk + '.branch7x7dbl_1', nxt=(c7, w71, (3, 0)), Hh=H),
                                                    spec(blk + '.branch_pool', out=out.slice(576, 768), pool_after=(3, 1, 1))],
                                            eps=eps, name=blk + '.1x1s')
            t = cb(t7, blk + '.branch7x7_2', pad=(0, 3), nxt=(192, w71, (3, 0)))
            cb(t, blk + '.branch7x7_3', pad=(3, 0), out=out.slice(192, 384))
            t = cb(td, blk + '.branch7x7dbl_2', pad=(3, 0), nxt=(c7, w17, (0, 3)))
            t = cb(t, blk + '.branch7x7dbl_3', pad=(0, 3), nxt=(c7, w71, (3, 0)))
            t = cb(t, blk + '.branch7x7dbl_4', pad=(3, 0), nxt=(192, w17, (0, 3)))
            cb(t, blk + '.branch7x7dbl_5', pad=(0, 3), out=out.slice(384, 576))
            x = out
            continue
        cb(x, blk + '.branch1x1', out=out.slice(0, 192))
        t = cb(x, blk + '.branch7x7_1', nxt=(c7, w17, (0, 3)))
        t = cb(t, blk + '.branch7x7_2', pad=(0, 3), nxt=(192, w71, (3, 0)))
        cb(t, blk + '.branch7x7_3', pad=(3, 0), out=out.slice(192, 384))
        t = cb(x, blk + '.branch7x7dbl_1', nxt=(c7, w71, (3, 0)))
        t = cb(t, blk + '.branch7x7dbl_2', pad=(3, 0), nxt=(c7, w17, (0, 3)))
        t = cb(t, blk + '.branch7x7dbl_3', pad=(0, 3), nxt=(c7, w71, (3, 0)))
        t = cb(t, blk + '.branch7x7dbl_4', pad=(3, 0), nxt=(192, w17, (0, 3)))
        cb(t, blk + '.branch7x7dbl_5', pad=(0, 3), out=out.slice(384, 576))
        cb(x, blk + '.branch_pool', out=out.slice(576, 768), pool_after=(3, 1, 1))
        x = out
    # AuxLogits (train mode only, inception.py:130-134, InceptionAux :361-395)
    if 'AuxLogits.conv0.conv.weight' in sd:
        t = tn.avgpool(x, 5, 3, 0)
        t = cb(t, 'AuxLogits.conv0')
        t = cb(t, 'AuxLogits.conv1')
        tn.head(t, sd, 'AuxLogits.fc', loss_weight=0.4, which='aux')
    blk = 'Mixed_7a'
    H2 = (x.H - 3) // 2 + 1
    out = tn.alloc(H2, H2, 1280)
    if fuse:
        t, t7 = tn.conv_bn_group(x, sd, [spec(blk + '.branch3x3_1'), spec(blk + '.branch7x7x3_1', nxt=(192, (1, 7), (0, 3)), Hh=x.H)],
                                 eps=eps, name=blk + '.1x1s')
    else:
        t = cb(x, blk + '.branch3x3_1')
        t7 = cb(x, blk + '.branch7x7x3_1', nxt=(192, (1, 7), (0, 3)))
    cb(t, blk + '.branch3x3_2', (2, 2), out=out.slice(0, 320))
    t = t7
    t = cb(t, blk + '.branch7x7x3_2', pad=(0, 3), nxt=(192, (7, 1), (3, 0)))
    t = cb(t, blk + '.branch7x7x3_3', pad=(3, 0))
    cb(t, blk + '.branch7x7x3_4', (2, 2), out=out.slice(320, 512))
    tn.maxpool(x, 3, 2, 0, out=out.slice(512, 1280))
    x = out
    for blk in ('Mixed_7b', 'Mixed_7c'):
        H = x.H
        out = tn.alloc(H, H, 2048)
        if fuse:
            _, t, td, _ = tn.conv_bn_group(x, sd, [spec(blk + '.branch1x1', out=out.slice(0, 320)),
                                                   spec(blk + '.branch3x3_1', nxt=(384, (3, 3), (1, 1)), Hh=H),
                                                   spec(blk + '.branch3x3dbl_1', nxt=(384, (3, 3), (1, 1)), Hh=H),
                                                   spec(blk + '.branch_pool', out=out.slice(1856, 2048), pool_after=(3, 1, 1))],
                                           eps=eps, name=blk + '.1x1s')
            cb(t, blk + '.branch3x3_2a', pad=(0, 1), out=out.slice(320, 704))
            cb(t, blk + '.branch3x3_2b', pad=(1, 0), out=out.slice(704, 1088))
            t = cb(td, blk + '.branch3x3dbl_2', pad=(1, 1), nxt=(384, (3, 3), (1, 1)))
            cb(t, blk + '.branch3x3dbl_3a', pad=(0, 1), out=out.slice(1088, 1472))
            cb(t, blk + '.branch3x3dbl_3b', pad=(1, 0), out=out.slice(1472, 1856))
            x = out
            continue
        cb(x, blk + '.branch1x1', out=out.slice(0, 320))
        t = cb(x, blk + '.branch3x3_1', nxt=(384, (3, 3), (1, 1)))        # read by the 1x3 and the 3x1 sibling
        cb(t, blk + '.branch3x3_2a', pad=(0, 1), out=out.slice(320, 704))
        cb(t, blk + '.branch3x3_2b', pad=(1, 0), out=out.slice(704, 1088))
        t = cb(x, blk + '.branch3x3dbl_1', nxt=(384, (3, 3), (1, 1)))
        t = cb(t, blk + '.branch3x3dbl_2', pad=(1, 1), nxt=(384, (3, 3), (1, 1)))
        cb(t, blk + '.branch3x3dbl_3a', pad=(0, 1), out=out.slice(1088, 1472))
        cb(t, blk + '.branch3x3dbl_3b', pad=(1, 0), out=out.slice(1472, 1856))
        cb(x, blk + '.branch_pool', out=out.slice(1856, 2048), pool_after=(3, 1, 1))
        x = out
    tn.head(x, sd, 'fc', dropout_p=0.5)


# ---- AlexNet / VGG (torchvision alexnet.py, vgg.py): conv(+bias)(+BN)+ReLU stacks, max pools, Dropout / Linear classifier ----
def _build_plain_train(tn, sd, arch):
    if arch == 'alexnet':
        seq, pool_k, pool_s = list(ALEXNET_CFG), 3, 2
    else:
        bn = arch.endswith('_bn')
        seq, idx = [], 0
        for v in VGG_CFG[arch[:-3] if bn else arch]:
            if v == 'M':
                seq.append('M'); idx += 1
            else:
                seq.append((idx, 3, 1, 1, idx + 1 if bn else None)); idx += 3 if bn else 2
        pool_k, pool_s = 2, 2
    x, first, H = None, True, tn.R
    for pos, item in enumerate(seq):
        if item == 'M':
            x = tn.maxpool(x, pool_k, pool_s, 0)
            H = x.H
            continue
        idx, k, stride, pad = item[:4]
        bn_idx = item[4] if len(item) > 4 else None
        conv = 'features.%d' % idx
        Co = int(sd[conv + '.weight'].shape[0])
        Ho = (H + 2 * pad - k) // stride + 1
        nxt = seq[pos + 1] if pos + 1 < len(seq) else 'M'
        opad = (0, 0)
        if nxt != 'M':
            nCo = int(sd['features.%d.weight' % nxt[0]].shape[0])
            opad = tn.border(Ho, Ho, Co, nCo, (nxt[1], nxt[1]), (nxt[2], nxt[2]), (nxt[3], nxt[3]))
        if bn_idx is not None:
            x = tn.conv_bn(x, sd, conv, 'features.%d' % bn_idx, (stride, stride), (pad, pad), out_pad=opad, stem=first, conv_bias=True)
        else:
            x = tn.conv_act(x, sd, conv, (stride, stride), (pad, pad), out_pad=opad, stem=first)
        first, H = False, Ho
    lin = sorted(int(k.split('.')[1]) for k in sd if k.startswith('classifier.') and k.endswith('.weight'))
    assert len(lin) == 3
    feat = int(sd['classifier.%d.weight' % lin[0]].shape[1])
    assert feat == x.C * H * H, '%s: a %d px input gives a %dx%d feature map; the classifier expects %d features (224 px)' % (
        arch, tn.R, H, H, feat)
    if arch == 'alexnet':                                            # Dropout, Linear, ReLU, Dropout, Linear, ReLU, Linear
        x = tn.drop(x, 0.5, 'classifier.0')
    w1 = sd['classifier.%d.weight' % lin[0]]
    x = tn.conv_act(x, sd, 'classifier.%d' % lin[0], weight=w1.view(int(w1.shape[0]), -1, H, H), linear=True)
    x = tn.drop(x, 0.5, 'classifier.drop1')
    w2 = sd['classifier.%d.weight' % lin[1]]
    x = tn.conv_act(x, sd, 'classifier.%d' % lin[1], weight=w2.view(int(w2.shape[0]), int(w2.shape[1]), 1, 1), linear=True)
    tn.head(x, sd, 'classifier.%d' % lin[2], dropout_p=0.0 if arch == 'alexnet' else 0.5)


# ---- SqueezeNet 1.1 (torchvision squeezenet.py; reference neuston_models.py:30-33 swaps classifier[1]) ----
def _build_squeezenet_train(tn, sd):
    x = tn.conv_act(None, sd, 'features.0', (2, 2), (0, 0), stem=True)
    for idx in (2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12):
        if idx in (2, 5, 8):
            x = tn.maxpool(x, 3, 2, 0, ceil_mode=True)
            continue
        pre = 'features.%d' % idx
        sq = int(sd[pre + '.squeeze.weight'].shape[0])
        e1, e3 = int(sd[pre + '.expand1x1.weight'].shape[0]), int(sd[pre + '.expand3x3.weight'].shape[0])
        t = tn.conv_act(x, sd, pre + '.squeeze', out_pad=tn.border(x.H, x.W, sq, e3, (3, 3), pad=(1, 1)))
        out = tn.alloc(x.H, x.W, e1 + e3)
        tn.conv_act(t, sd, pre + '.expand1x1', out=out.slice(0, e1))
        tn.conv_act(t, sd, pre + '.expand3x3', pad=(1, 1), out=out.slice(e1, e1 + e3))
        x = out
    x = tn.drop(x, 0.5, 'classifier.0')
    n_classes = int(sd['classifier.1.weight'].shape[0])
    y = tn.conv_act(x, sd, 'classifier.1', co_pad=(n_classes + 7) // 8 * 8)
    tn.head(y, None, None, identity_classes=n_classes)


# ---- DenseNet (torchvision densenet.py): pre-activation dense layers over a growing concatenation ----
def _build_densenet_train(tn, sd):
    blocks, b = [], 1
    while ('features.denseblock%d.denselayer1.conv1.weight' % b) in sd:
        n = 1
        while ('features.denseblock%d.denselayer%d.conv1.weight' % (b, n + 1)) in sd:
            n += 1
        blocks.append(n)
        b += 1
    growth = int(sd['features.denseblock1.denselayer1.conv2.weight'].shape[0])
    C_in = int(sd['features.conv0.weight'].shape[0])
    a = tn.conv_bn(None, sd, 'features.conv0', 'features.norm0', (2, 2), (3, 3), stem=True)
    H = (a.H + 2 - 3) // 2 + 1
    buf = tn.alloc(H, H, C_in + blocks[0] * growth)
    tn.maxpool(a, 3, 2, 1, out=buf.slice(0, C_in))
    for bi, n_layers in enumerate(blocks):
        for li in range(n_layers):
            pre = 'features.denseblock%d.denselayer%d' % (bi + 1, li + 1)
            t1 = tn.bn_act(buf.slice(0, C_in), sd, pre + '.norm1')
            mid = int(sd[pre + '.conv1.weight'].shape[0])
            t2 = tn.conv_bn(t1, sd, pre + '.conv1', pre + '.norm2', out_pad=tn.border(H, H, mid, growth, (3, 3), pad=(1, 1)))
            tn.conv_act(t2, sd, pre + '.conv2', pad=(1, 1), relu=False, bias=False, out=buf.slice(C_in, C_in + growth))
            C_in += growth
        if bi + 1 < len(blocks):                       # transition: norm -> relu -> conv 1x1 -> avg_pool 2x2
            pre = 'features.transition%d' % (bi + 1)
            t1 = tn.bn_act(buf, sd, pre + '.norm', first=True)
            t2 = tn.conv_act(t1, sd, pre + '.conv', relu=False, bias=False)
            Co = t2.C
            H = (H - 2) // 2 + 1
            nbuf = tn.alloc(H, H, Co + blocks[bi + 1] * growth)
            tn.avgpool(t2, 2, 2, 0, out=nbuf.slice(0, Co))
            buf, C_in = nbuf, Co
    t = tn.bn_act(buf, sd, 'features.norm5', first=True)
    tn.head(t, sd, 'classifier')
