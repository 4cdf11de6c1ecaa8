"""TRAIN-step primitives against torch autograd (fp32, CPU/GPU) -- tolerances stated per test.
Everything goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# torch references below are fp32: no TF32 in cuDNN / cuBLAS
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

WGRAD_CASES = [
    # name, B, Cin, H, W, Cout, kh, kw, stride, pad
    ('3x3p1_64_64', 3, 64, 14, 14, 64, 3, 3, (1, 1), (1, 1)),
    ('1x1_256_64', 2, 256, 9, 9, 64, 1, 1, (1, 1), (0, 0)),
    ('3x3s2p1_64_128', 2, 64, 28, 28, 128, 3, 3, (2, 2), (1, 1)),
    ('1x1s2_64_256', 2, 64, 15, 15, 256, 1, 1, (2, 2), (0, 0)),
    ('3x3p1_128_192', 5, 128, 17, 17, 192, 3, 3, (1, 1), (1, 1)),
    ('1x7_160_160', 2, 160, 17, 17, 160, 1, 7, (1, 1), (0, 3)),
    ('7x7s2p3_8_64', 2, 8, 32, 32, 64, 7, 7, (2, 2), (3, 3)),
    # Cin <= 32: 32-channel x blocks (64-byte rows), four per M = 128
    ('3x3_32_32', 3, 32, 21, 21, 32, 3, 3, (1, 1), (0, 0)),
    ('3x3p1_32_64', 2, 32, 19, 19, 64, 3, 3, (1, 1), (1, 1)),
    ('1x1_32_80', 4, 32, 11, 13, 80, 1, 1, (1, 1), (0, 0)),
    ('3x3s2_24_48', 2, 24, 23, 23, 48, 3, 3, (2, 2), (0, 0)),
    ('5x5p2_16_32', 2, 16, 12, 12, 32, 5, 5, (1, 1), (2, 2)),
    # stride 1 on feature maps that 128-pixel rectangles tile well: the WINDOW variant (one input patch per tile, taps by
    # descriptor shift); ragged right / bottom tiles, two channel blocks, 16- and 64-pixel-wide tiles, an odd block count
    ('3x3_80_192_win', 2, 80, 34, 34, 192, 3, 3, (1, 1), (0, 0)),
    ('3x3p1_32_64_win', 2, 32, 47, 47, 64, 3, 3, (1, 1), (1, 1)),
    ('1x7_64_96_win', 1, 64, 30, 62, 96, 1, 7, (1, 1), (0, 3)),
    ('5x5p2_48_64_win', 2, 48, 31, 31, 64, 5, 5, (1, 1), (2, 2)),
    ('3x3p1_128_256_win', 1, 128, 28, 28, 256, 3, 3, (1, 1), (1, 1)),
]


@pytest.mark.parametrize('dt', ['bf16', 'fp16'])
@pytest.mark.parametrize('case', WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_conv_wgrad(cuda, case, dt, monkeypatch):
    from ifcb_classifier_b200 import _lib
    name, B, Cin, H, W, Cout, kh, kw, stride, pad = case
    if name.endswith('_win'):            # the library keeps the WINDOW variant for maps >= 64 x 64 unless forced
        monkeypatch.setenv('IFCB_WGRAD_WINDOW', '2')
    tdt = torch.bfloat16 if dt == 'bf16' else torch.float16
    g = torch.Generator().manual_seed(sum(name.encode()))
    x = torch.randn(B, Cin, H, W, generator=g).to(tdt).float()
    P = (H + 2 * pad[0] - kh) // stride[0] + 1
    Q = (W + 2 * pad[1] - kw) // stride[1] + 1
    dz = (torch.randn(B, Cout, P, Q, generator=g) * 0.1).to(tdt).float()
    # reference: autograd of conv2d wrt the weight, fp32 on the same 16-bit-rounded operands
    w = torch.zeros(Cout, Cin, kh, kw, requires_grad=True)
    F.conv2d(x, w, stride=stride, padding=pad).backward(dz)
    want = w.grad.permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin)          # [co, tap, ci]
    xd = x.permute(0, 2, 3, 1).contiguous().to(cuda).to(tdt)
    dzd = dz.permute(0, 2, 3, 1).contiguous().to(cuda).to(tdt)
    dw = torch.zeros(Cout, kh * kw, Cin, dtype=torch.float32, device=cuda)
    d = _lib.WgradDesc()
    d.d_in, d.in_ld, d.Cin = xd.data_ptr(), Cin, Cin
    d.batch, d.H, d.W = B, H, W
    d.kh, d.kw, d.stride_h, d.stride_w, d.pad_h, d.pad_w = kh, kw, stride[0], stride[1], pad[0], pad[1]
    d.d_dout, d.dout_ld, d.Cout = dzd.data_ptr(), Cout, Cout
    d.d_dweight = dw.data_ptr()
    d.dtype = _lib.IFCB_ACT_BF16 if dt == 'bf16' else _lib.IFCB_ACT_FP16
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().ifcb_conv_wgrad(C.byref(d), stream), 'conv_wgrad')
    torch.cuda.synchronize()
    got = dw.cpu()
    # exact products of 16-bit operands, fp32 accumulation in a different order: 1e-3 of the largest entry
    err = float((got - want).abs().max())
    assert err <= 1e-3 * float(want.abs().max()) + 1e-5, (name, err, float(want.abs().max()))


# ------------------------------------------------------------------------------------------------
# streaming TRAIN kernels (csrc/train_ops.cu) against torch autograd, fp32 on the same 16-bit inputs
# ------------------------------------------------------------------------------------------------
def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _mk(cuda, B, H, W, Cc, tdt, pad=(0, 0), ld=None, c0=0, gen=None, scale=1.0, fill=True):
    """Random NHWC view (optionally a channel slice of a wider, bordered tensor) + its fp32 NCHW value."""
    from ifcb_classifier_b200.graph import View
    ld = ld or Cc
    t = torch.zeros(B, H + 2 * pad[0], W + 2 * pad[1], ld, dtype=tdt, device=cuda)
    v = View(t, c0, c0 + Cc, pad)
    if fill:
        val = (torch.randn(B, H, W, Cc, generator=gen) * scale).to(tdt)
        v.interior().copy_(val.to(cuda))
        return v, val.float().permute(0, 3, 1, 2).contiguous().to(cuda)
    return v, None


def _read(v):
    return v.interior().float().permute(0, 3, 1, 2).contiguous()


BN_CASES = [
    # name, B, H, W, C, relu, residual, in/out pad, ld, c0
    ('c64_relu', 4, 9, 9, 64, True, False, (0, 0), None, 0),
    ('c80_slice_pad', 3, 7, 5, 80, True, False, (1, 1), 160, 32),
    ('c256_res_relu', 2, 6, 6, 256, True, True, (0, 0), None, 0),
    ('c128_norelu', 5, 4, 4, 128, False, False, (0, 0), None, 0),
    ('c2048', 2, 3, 3, 2048, True, False, (0, 0), None, 0),
    ('c48_big', 8, 35, 35, 48, True, False, (2, 2), None, 0),
]


@pytest.mark.parametrize('dt', ['bf16', 'fp16'])
@pytest.mark.parametrize('case', BN_CASES, ids=[c[0] for c in BN_CASES])
def test_bn_train_fwd_bwd(cuda, case, dt):
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.train import _vd
    name, B, H, W, Cc, relu, has_res, pad, ld, c0 = case
    tdt = torch.bfloat16 if dt == 'bf16' else torch.float16
    cdt = _lib.IFCB_ACT_BF16 if dt == 'bf16' else _lib.IFCB_ACT_FP16
    g = torch.Generator().manual_seed(sum(name.encode()))
    L = _lib.lib()
    z, z32 = _mk(cuda, B, H, W, Cc, tdt, gen=g, scale=2.0)
    z32 = z32 + 0.0
    out, _ = _mk(cuda, B, H, W, Cc, tdt, pad=pad, ld=ld, c0=c0, fill=False)
    res, res32 = _mk(cuda, B, H, W, Cc, tdt, pad=pad, gen=g) if has_res else (None, None)
    gamma = (torch.rand(Cc, generator=g) + 0.5).to(cuda)
    beta = (torch.randn(Cc, generator=g) * 0.2).to(cuda)
    rm, rv = torch.zeros(Cc, device=cuda), torch.ones(Cc, device=cuda)
    mean, invstd = torch.zeros(Cc, device=cuda), torch.zeros(Cc, device=cuda)
    acc = torch.zeros(8192, dtype=torch.float64, device=cuda)       # 64 KB scratch: accumulators + coefficient floats
    zd, od = _vd(z), _vd(out)
    rd = _vd(res) if has_res else None
    _lib.check(L.ifcb_bn_stats(C.byref(zd), B, cdt, 1e-3, 0.1, acc.data_ptr(), mean.data_ptr(), invstd.data_ptr(), rm.data_ptr(),
                               rv.data_ptr(), _stream()), 'bn_stats')
    _lib.check(L.ifcb_bn_apply(C.byref(zd), C.byref(od), C.byref(rd) if rd is not None else None, B, cdt, mean.data_ptr(), invstd.data_ptr(),
                               gamma.data_ptr(), beta.data_ptr(), 1 if relu else 0, _stream()), 'bn_apply')
    # reference
    zr = z32.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rr = res32.clone().requires_grad_(True) if has_res else None
    rm_r, rv_r = torch.zeros(Cc, device=cuda), torch.ones(Cc, device=cuda)
    y = F.batch_norm(zr, rm_r, rv_r, gr, br, training=True, momentum=0.1, eps=1e-3)
    if has_res:
        y = y + rr
    if relu:
        y = F.relu(y)
    got = _read(out)
    tol = 2.0 ** (-8 if dt == 'bf16' else -11)               # one 16-bit rounding of the output
    assert float((got - y.detach()).abs().max()) <= tol * (1.0 + float(y.detach().abs().max())), name
    assert float(acc[:4096].abs().max()) == 0.0               # accumulators cleared for the next layer
    assert torch.allclose(rm, rm_r, rtol=1e-4, atol=1e-5) and torch.allclose(rv, rv_r, rtol=1e-4, atol=1e-5)
    # backward: dy arrives as a 16-bit tensor; mask from OUR forward output so both sides agree on it
    dy, dy32 = _mk(cuda, B, H, W, Cc, tdt, gen=g, scale=0.5)
    dres, _ = _mk(cuda, B, H, W, Cc, tdt, gen=g) if has_res else (None, None)
    dres0 = _read(dres) if has_res else None
    dyd = _vd(dy)
    dgam, dbet = torch.zeros(Cc, device=cuda), torch.zeros(Cc, device=cuda)
    y.backward(dy32)
    _lib.check(L.ifcb_bn_backward(C.byref(dyd), C.byref(od) if (relu and has_res) else None, C.byref(zd), C.byref(dyd),
                                  C.byref(_vd(dres)) if has_res else None, 1, 1 if relu else 0, B, cdt, mean.data_ptr(), invstd.data_ptr(),
                                  gamma.data_ptr(), beta.data_ptr(), acc.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), _stream()), 'bn_backward')
    torch.cuda.synchronize()
    dz = _read(dy)
    scale = float(zr.grad.abs().max())
    # (an element whose pre-activation is within fp32 noise of zero may land on the other side of the ReLU)
    bad = float(((dz - zr.grad).abs() > 2 * tol * scale + 1e-6).float().mean())
    assert bad <= 1e-4, (name, bad, float((dz - zr.grad).abs().max()), scale)
    assert torch.allclose(dgam, gr.grad, rtol=2e-3, atol=2e-3 * float(gr.grad.abs().max()))
    assert torch.allclose(dbet, br.grad, rtol=2e-3, atol=2e-3 * float(br.grad.abs().max()))
    if has_res:
        want = dres0 + rr.grad
        assert float((_read(dres) - want).abs().max()) <= 2 * tol * float(want.abs().max())


POOL_CASES = [
    # name, kind, B, H, W, C, k, s, p, in pad, out (ld, c0)
    ('max3s2', 'max', 3, 15, 15, 64, 3, 2, 0, (0, 0), (64, 0)),
    ('max3s2p1', 'max', 2, 14, 14, 64, 3, 2, 1, (0, 0), (64, 0)),
    ('max3s2_slice', 'max', 2, 17, 17, 96, 3, 2, 0, (1, 1), (192, 96)),
    ('avg3s1p1', 'avg', 2, 9, 9, 192, 3, 1, 1, (1, 1), (192, 0)),
    ('avg5s3', 'avg', 3, 17, 17, 768, 5, 3, 0, (0, 0), (768, 0)),
]


@pytest.mark.parametrize('dt', ['bf16', 'fp16'])
@pytest.mark.parametrize('case', POOL_CASES, ids=[c[0] for c in POOL_CASES])
def test_pool_train_fwd_bwd(cuda, case, dt):
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.train import _vd
    name, kind, B, H, W, Cc, k, s, p, ipad, (old, oc0) = case
    tdt = torch.bfloat16 if dt == 'bf16' else torch.float16
    cdt = _lib.IFCB_ACT_BF16 if dt == 'bf16' else _lib.IFCB_ACT_FP16
    g = torch.Generator().manual_seed(sum(name.encode()))
    L = _lib.lib()
    x, x32 = _mk(cuda, B, H, W, Cc, tdt, pad=ipad, gen=g)
    # post-ReLU-like input: many exact ties at zero exercise torch's first-maximum rule
    x.interior().clamp_(min=0)
    x32 = x32.clamp(min=0)
    P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    y, _ = _mk(cuda, B, P, Q, Cc, tdt, ld=old, c0=oc0, fill=False)
    xr = x32.clone().requires_grad_(True)
    xd, yd = _vd(x), _vd(y)
    if kind == 'max':
        idx = torch.zeros(B, P, Q, Cc, dtype=torch.uint8, device=cuda)
        _lib.check(L.ifcb_maxpool_fwd_train(C.byref(xd), C.byref(yd), idx.data_ptr(), B, k, s, p, cdt, _stream()), 'maxpool_fwd')
        yr = F.max_pool2d(xr, k, s, p)
    else:
        _lib.check(L.ifcb_avgpool_fwd(C.byref(xd), C.byref(yd), B, k, s, p, cdt, _stream()), 'avgpool_fwd')
        yr = F.avg_pool2d(xr, k, s, p)
    tol = 2.0 ** (-8 if dt == 'bf16' else -11)
    assert float((_read(y) - yr.detach()).abs().max()) <= (tol * float(yr.detach().abs().max()) if kind == 'avg' else 0.0)
    dy, dy32 = _mk(cuda, B, P, Q, Cc, tdt, ld=old, c0=oc0, gen=g)
    yr.backward(dy32)
    for accumulate in (0, 1):
        dx, dx0 = _mk(cuda, B, H, W, Cc, tdt, gen=g)
        dyd, dxd = _vd(dy), _vd(dx)
        if kind == 'max':
            _lib.check(L.ifcb_maxpool_bwd(C.byref(dyd), idx.data_ptr(), C.byref(dxd), accumulate, B, k, s, p, cdt, _stream()), 'maxpool_bwd')
        else:
            _lib.check(L.ifcb_avgpool_bwd(C.byref(dyd), C.byref(dxd), accumulate, B, k, s, p, cdt, _stream()), 'avgpool_bwd')
        want = xr.grad + (dx0 if accumulate else 0)
        err = float((_read(dx) - want).abs().max())
        assert err <= 2 * tol * float(want.abs().max()), (name, accumulate, err)


def test_dilate_and_nchw_to_nhwc(cuda):
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.train import _vd
    L = _lib.lib()
    g = torch.Generator().manual_seed(5)
    x, x32 = _mk(cuda, 2, 5, 4, 32, torch.bfloat16, gen=g)
    out, _ = _mk(cuda, 2, 10, 8, 32, torch.bfloat16, fill=False)
    _lib.check(L.ifcb_dilate(C.byref(_vd(x)), C.byref(_vd(out)), 2, 2, 2, _stream()), 'dilate')
    want = torch.zeros(2, 32, 10, 8, device=cuda)
    want[:, :, 0:9:2, 0:7:2] = x32
    assert torch.equal(_read(out), want)
    img = torch.randn(3, 3, 11, 13, generator=g).to(cuda)
    o8, _ = _mk(cuda, 3, 11, 13, 8, torch.bfloat16, fill=False)
    _lib.check(L.ifcb_nchw_to_nhwc(img.data_ptr(), 3, C.byref(_vd(o8)), 3, _lib.IFCB_ACT_BF16, _stream()), 'nchw_to_nhwc')
    got = _read(o8)
    assert torch.equal(got[:, :3], img.to(torch.bfloat16).float()) and float(got[:, 3:].abs().max()) == 0.0


@pytest.mark.parametrize('drop', [False, True])
def test_head_train_fwd_bwd(cuda, drop):
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.train import _vd
    L = _lib.lib()
    B, H, W, Cc, K = 6, 3, 3, 512, 37
    g = torch.Generator().manual_seed(11)
    x, x32 = _mk(cuda, B, H, W, Cc, torch.bfloat16, gen=g)
    Wt = (torch.randn(K, Cc, generator=g) * 0.05).to(cuda)
    bias = (torch.randn(K, generator=g) * 0.1).to(cuda)
    labels = torch.randint(0, K, (B,), generator=g).to(cuda)
    ds = None
    if drop:
        ds = torch.zeros(B * Cc, device=cuda)
        _lib.check(L.ifcb_dropout_scale(ds.data_ptr(), B * Cc, 0.5, 1234, _stream()), 'dropout_scale')
        torch.cuda.synchronize()
        frac = float((ds == 0).float().mean())
        assert 0.4 < frac < 0.6 and set(ds.unique().tolist()) == {0.0, 2.0}
    pooled, logits, dlogits = torch.zeros(B, Cc, device=cuda), torch.zeros(B, K, device=cuda), torch.zeros(B, K, device=cuda)
    loss = torch.zeros(2, device=cuda)
    _lib.check(L.ifcb_head_train_fwd(C.byref(_vd(x)), B, _lib.IFCB_ACT_BF16, ds.data_ptr() if drop else None, Wt.data_ptr(), bias.data_ptr(),
                                     labels.data_ptr(), K, 0.4, pooled.data_ptr(), logits.data_ptr(), dlogits.data_ptr(), loss.data_ptr(),
                                     _stream()), 'head_train_fwd')
    xr = x32.clone().requires_grad_(True)
    Wr, br = Wt.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    pr = xr.mean((2, 3))
    if drop:
        pr = pr * ds.view(B, Cc)
    lr_ = F.linear(pr, Wr, br)
    lref = 0.4 * F.cross_entropy(lr_, labels)
    lref.backward()
    assert torch.allclose(logits, lr_.detach(), rtol=1e-4, atol=1e-5)
    assert abs(float(loss[0]) - float(lref)) <= 1e-5 * max(1.0, abs(float(lref)))
    dW, db = torch.zeros(K, Cc, device=cuda), torch.zeros(K, device=cuda)
    dx, dx0 = _mk(cuda, B, H, W, Cc, torch.bfloat16, gen=g, scale=1e-3)
    _lib.check(L.ifcb_head_bwd(C.byref(_vd(dx)), 1, B, _lib.IFCB_ACT_BF16, ds.data_ptr() if drop else None, Wt.data_ptr(), pooled.data_ptr(),
                               dlogits.data_ptr(), K, dW.data_ptr(), db.data_ptr(), _stream()), 'head_bwd')
    torch.cuda.synchronize()
    assert torch.allclose(dW, Wr.grad, rtol=1e-4, atol=1e-6) and torch.allclose(db, br.grad, rtol=1e-4, atol=1e-6)
    want = dx0 + xr.grad
    assert float((_read(dx) - want).abs().max()) <= 2.0 ** -7 * float(want.abs().max())


def test_adam_matches_torch(cuda):
    from ifcb_classifier_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(3)
    n = 10007
    w0 = torch.randn(n, generator=g).to(cuda)
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.Adam([p], lr=1e-3)
    w, m, v = w0.clone(), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    for step in range(1, 6):
        gr = (torch.randn(n, generator=g) * 10 ** float(torch.randint(-4, 1, (1,), generator=g))).to(cuda)
        p.grad = gr.clone()
        opt.step()
        _lib.check(L.ifcb_adam_step(w.data_ptr(), (gr * 4).data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, step, 0.25,
                                    _stream()), 'adam')
        torch.cuda.synchronize()
        assert float((w - p.detach()).abs().max()) <= 2e-6, step


def test_conv_repack_layouts(cuda):
    from ifcb_classifier_b200 import _lib
    L = _lib.lib()
    Co, Ci, kh, kw = 48, 80, 3, 3
    g = torch.Generator().manual_seed(2)
    w = torch.randn(Co, Ci, kh, kw, generator=g)
    master = w.permute(0, 2, 3, 1).reshape(Co, kh * kw, Ci).contiguous().to(cuda)
    gf, gd = _lib.conv_geometry(Ci, Co, kh, kw), _lib.conv_geometry(Co, Ci, kh, kw)
    wf = torch.zeros(gf['Cout_pad'], kh * kw * gf['Cin_pad'], dtype=torch.bfloat16, device=cuda)
    wd = torch.zeros(gd['Cout_pad'], kh * kw * gd['Cin_pad'], dtype=torch.bfloat16, device=cuda)
    _lib.check(L.ifcb_conv_repack(master.data_ptr(), Co, kh * kw, Ci, wf.data_ptr(), gf['Cin_pad'], wd.data_ptr(), gd['Cin_pad'],
                                  _lib.IFCB_ACT_BF16, _stream()), 'repack')
    torch.cuda.synchronize()
    wb = w.to(torch.bfloat16).float()
    got_f = wf.float().view(gf['Cout_pad'], kh * kw, gf['Cin_pad'])[:Co, :, :Ci].cpu()
    assert torch.equal(got_f, wb.permute(0, 2, 3, 1).reshape(Co, kh * kw, Ci))
    # data-gradient operand = packed weights of conv_transpose: [ci][flipped tap][co]
    wt = wb.flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, kh * kw, Co)
    got_d = wd.float().view(gd['Cout_pad'], kh * kw, gd['Cin_pad'])[:Ci, :, :Co].cpu()
    assert torch.equal(got_d, wt)


# ------------------------------------------------------------------------------------------------
# whole TRAIN step against the oracle (torchvision fp32 autograd + Adam, oracle/train_ref.py)
# ------------------------------------------------------------------------------------------------
DGRAD_CASES = [
    # name, B, Cin, H, W, Cout, kh, kw, stride, pad
    ('3x3p1_64_96', 3, 64, 14, 14, 96, 3, 3, (1, 1), (1, 1)),
    ('1x1_256_64', 2, 256, 9, 9, 64, 1, 1, (1, 1), (0, 0)),
    ('3x3s2p1_64_128', 2, 64, 28, 28, 128, 3, 3, (2, 2), (1, 1)),
    ('3x3s2p0_288_384', 2, 288, 35, 35, 384, 3, 3, (2, 2), (0, 0)),
    ('1x1s2_64_256', 2, 64, 15, 15, 256, 1, 1, (2, 2), (0, 0)),
    ('1x7_160_192', 2, 160, 17, 17, 192, 1, 7, (1, 1), (0, 3)),
    ('5x5p2_48_64', 2, 48, 35, 35, 64, 5, 5, (1, 1), (2, 2)),
    ('3x3p1_512_512_2x2', 4, 512, 2, 2, 512, 3, 3, (1, 1), (1, 1)),
    ('5x5_128_768_to1x1', 4, 128, 5, 5, 768, 5, 5, (1, 1), (0, 0)),
]


@pytest.mark.parametrize('accumulate', [False, True])
@pytest.mark.parametrize('case', DGRAD_CASES, ids=[c[0] for c in DGRAD_CASES])
def test_conv_dgrad(cuda, case, accumulate):
    """Data gradient = forward tcgen05 kernel + repacked (flipped / transposed) filter (+ dilation)."""
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.graph import PlanBuilder, View
    from ifcb_classifier_b200.train import build_dgrad
    name, B, Cin, H, W, Cout, kh, kw, stride, pad = case
    tdt = torch.bfloat16
    g = torch.Generator().manual_seed(sum(name.encode()))
    P = (H + 2 * pad[0] - kh) // stride[0] + 1
    Q = (W + 2 * pad[1] - kw) // stride[1] + 1
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) * 0.05).to(tdt).float()
    dy, dy32 = _mk(cuda, B, P, Q, Cout, tdt, gen=g)
    dx, dx0 = _mk(cuda, B, H, W, Cin, tdt, gen=g)
    bp = PlanBuilder(B, cuda, 'bf16')
    dg = build_dgrad(bp, dy, dx, Cout, Cin, kh, kw, stride, pad, accumulate)
    master = w.permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin).contiguous().to(cuda)
    _lib.check(_lib.lib().ifcb_conv_repack(master.data_ptr(), Cout, kh * kw, Cin, None, 0, dg['weight'].data_ptr(), dg['Cin_pad'],
                                           _lib.IFCB_ACT_BF16, _stream()), 'repack')
    for r in dg['run']:
        r()
    torch.cuda.synchronize()
    # float64 reference: cuDNN's fp32 backward-data algorithms are themselves off by > 1 bf16 ulp on some shapes
    xr = torch.zeros(B, Cin, H, W, device=cuda, dtype=torch.float64, requires_grad=True)
    F.conv2d(xr, w.to(cuda).double(), stride=stride, padding=pad).backward(dy32.double())
    want = xr.grad.float() + (dx0 if accumulate else 0)
    err = float((_read(dx) - want).abs().max())
    assert err <= 2.0 ** -7 * float(want.abs().max()), (name, err, float(want.abs().max()))


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


def _cos(a, b):
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-20))


STEP_CASES = [('resnet18', 16, 96), ('resnet50', 8, 96), ('inception_v3', 8, 299)]


@pytest.mark.parametrize('case', STEP_CASES, ids=[c[0] for c in STEP_CASES])
def test_forward_backward_vs_oracle(cuda, case):
    """One training_step (bf16 storage, fp32 accumulate; dropout off on both sides).

    (1) Kernel parity, teacher-forced (tests/train_local.py): every unit of the step recomputed with
        torch fp32 ops from the tensors the path stored -- conv outputs, BN batch statistics, activations,
        dz, dW, dgamma, dbeta, fc gradients, and every activation gradient as the sum of its consumers'
        contributions -- within 2^-7 of the tensor's largest entry (a few bf16 roundings).
    (2) End to end against the oracle (fp32 autograd of the same torchvision module,
        oracle/train_ref.py): |loss - ref| <= 2e-2 * ref.  Gradients are REPORTED, not gated, beyond
        resnet18: a deep ReLU+BN network at random init amplifies one-ulp activation differences
        layer by layer, so bf16 storage and fp32 disagree on ReLU masks; the same disagreement shows
        between the oracle and the oracle with bf16 storage rounding inserted (printed as 'rounded vs fp32')."""
    import copy
    from oracle import train_ref
    from tests.fixtures import ref_model
    from tests.train_local import local_parity
    from ifcb_classifier_b200.train import TrainNet
    arch, B, R = case
    n_classes = 10
    model = ref_model(arch, n_classes, seed=1).to(cuda)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(B, 3, R, R, generator=g).to(cuda)
    y = torch.randint(0, n_classes, (B,), generator=g).to(cuda)
    qmodel = copy.deepcopy(model)
    train_ref.with_storage_rounding(qmodel, torch.bfloat16)
    model.load_state_dict(qmodel.state_dict())                 # conv weights representable in bf16 on every side
    net = TrainNet(arch, model.state_dict(), B, device=cuda, dtype='bf16', dropout=False, R=R, keep_dy=True)
    loss = float(net.forward_backward(x, y))
    ok, stats = local_parity(net, (lambda nm: 1e-3) if arch == 'inception_v3' else (lambda nm: 1e-5))
    worst = sorted(stats, key=lambda t: -t[1])[:5]
    print('%s teacher-forced: %d checks, failing %s, largest rel L2 %s' %
          (arch, len(stats), [t for t in stats if not t[3]][:8], [(t[0], round(t[1], 5)) for t in worst]))
    assert ok, [t for t in stats if not t[3]][:20]
    grads = {k: v.to(cuda).float() for k, v in net.grad_dict().items()}
    # a SECOND step on new data after an Adam update (operands repacked, scratch buffers and gradient
    # tensors reused): the same teacher-forced check must still hold
    net.adam()
    x2 = torch.rand(B, 3, R, R, generator=g).to(cuda)
    y2 = torch.randint(0, n_classes, (B,), generator=g).to(cuda)
    net.forward_backward(x2, y2)
    ok2, stats2 = local_parity(net, (lambda nm: 1e-3) if arch == 'inception_v3' else (lambda nm: 1e-5))
    assert ok2, ('second step', [t for t in stats2 if not t[3]][:20])
    net.inp.copy_(x)

    q_loss, q_grads = train_ref.forward_backward(qmodel, x, y, dropout=False)
    f_loss, f_grads = train_ref.forward_backward(model, x, y, dropout=False)
    assert sorted(grads) == sorted(f_grads)

    def cosine(ga, gb):
        dot = sum((ga[k] * gb[k]).sum() for k in gb)
        na = torch.sqrt(sum((v ** 2).sum() for v in ga.values()))
        nb = torch.sqrt(sum((v ** 2).sum() for v in gb.values()))
        return float(dot / (na * nb))

    print('%s end to end: loss %.5f, fp32 oracle %.5f, rounded oracle %.5f; gradient cosine ours vs fp32 %.4f, ours vs rounded %.4f, '
          'rounded vs fp32 %.4f' % (arch, loss, float(f_loss), float(q_loss), cosine(grads, f_grads), cosine(grads, q_grads),
                                   cosine(q_grads, f_grads)))
    assert abs(loss - float(f_loss)) <= 2e-2 * abs(float(f_loss)), (loss, float(f_loss))
    if arch == 'resnet18':
        assert cosine(grads, f_grads) >= 0.9


def test_train_steps_follow_oracle(cuda):
    """Ten Adam steps of resnet18 on separable synthetic classes: the first loss matches the oracle's
    within 1 %, both runs converge, running statistics stay within 5e-2 / 2e-2."""
    from oracle import train_ref
    from tests.fixtures import ref_model, class_rois
    from ifcb_classifier_b200.train import TrainNet
    import numpy as np
    B, R, n_classes = 32, 64, 4
    imgs, labels = class_rois(5 * B, n_classes, seed=3, hw=(R, R))
    X = torch.from_numpy(np.stack(imgs)).float().div(255)[:, None].repeat(1, 3, 1, 1)
    model = ref_model('resnet18', n_classes, seed=2).to(cuda)
    net = TrainNet('resnet18', model.state_dict(), B, device=cuda, dtype='bf16', R=R)
    batches = [(X[i * B:(i + 1) * B].to(cuda), labels[i * B:(i + 1) * B].to(cuda)) for i in range(5)] * 2
    first = float(net.step(*batches[0]))
    sd1 = net.state_dict()
    ours = [first] + [float(net.step(xb, yb)) for xb, yb in batches[1:]]
    import copy
    m1 = copy.deepcopy(model)
    train_ref.train_steps(m1, batches[:1])
    rsd1 = m1.state_dict()
    for k in ('bn1.running_mean', 'bn1.running_var', 'layer2.0.bn1.running_var', 'layer4.1.bn2.running_mean'):
        assert torch.allclose(sd1[k], rsd1[k].cpu(), rtol=2e-2, atol=5e-3), k      # BN statistics of the first batch
    ref = train_ref.train_steps(model, batches)
    print('ours', ours, 'ref', ref)
    # the first updates follow the oracle closely; later the two runs are different (equally valid)
    # realisations of a chaotic trajectory -- both must converge
    assert abs(ours[0] - ref[0]) <= 0.01 * abs(ref[0]), (ours, ref)      # (Adam's first update is ~ lr * sign(g): every later loss
    #                                                                        depends on the sign of each near-zero gradient entry)
    assert max(ours[-3:]) < 0.1 * ours[0] and max(ref[-3:]) < 0.1 * ref[0]
    sd, rsd = net.state_dict(), model.state_dict()
    assert list(sd.keys()) == list(rsd.keys())
    assert int(sd['bn1.num_batches_tracked']) == 10


FOLLOW_CASES = [('resnet50', 16, 96, 16), ('inception_v3', 8, 299, 16)]


@pytest.mark.parametrize('arch,B,R,steps', FOLLOW_CASES, ids=[c[0] for c in FOLLOW_CASES])
def test_baseline_train_models_converge_like_the_oracle(cuda, arch, B, R, steps):
    """The two BASELINE TRAIN models, Adam steps on separable synthetic classes beside the oracle (fp32
    torch autograd + torch.optim.Adam on the same batches, oracle/train_ref.py): the first loss agrees within 2 % (Inception: CE +
    0.4 CE_aux), both runs bring the loss down (mean of the last four losses < 0.6 x the first), and ours ends no worse than 2 x the oracle's
    final loss + 0.3 (measured: resnet50 0.19 vs 0.38, inception_v3 0.28 vs 0.12; Inception's first ten losses track the oracle's to ~5 %).  (Step-by-step
    equality is not expected: Adam's first update is ~ lr * sign(g), so the two runs are different realisations of a chaotic
    trajectory -- see test_train_steps_follow_oracle.)"""
    from oracle import train_ref
    from tests.fixtures import ref_model, class_rois
    from ifcb_classifier_b200.train import TrainNet
    import numpy as np
    n_classes, n_batches = 4, 4
    imgs, labels = class_rois(n_batches * B, n_classes, seed=5, hw=(R, R))
    X = torch.from_numpy(np.stack(imgs)).float().div(255)[:, None].repeat(1, 3, 1, 1)
    model = ref_model(arch, n_classes, seed=2).to(cuda)
    net = TrainNet(arch, model.state_dict(), B, device=cuda, dtype='bf16', R=R, dropout=False)      # dropout streams differ by design
    batches = [(X[i * B:(i + 1) * B].to(cuda), labels[i * B:(i + 1) * B].to(cuda)) for i in range(n_batches)] * (steps // n_batches)
    ours = [float(net.step(xb, yb)) for xb, yb in batches]
    ref = train_ref.train_steps(model, batches, dropout=False)
    print(arch, 'ours', [round(v, 3) for v in ours], 'ref', [round(v, 3) for v in ref])
    assert abs(ours[0] - ref[0]) <= 0.02 * abs(ref[0]), (ours[0], ref[0])
    o_end, r_end = sum(ours[-4:]) / 4, sum(ref[-4:]) / 4
    assert r_end < 0.6 * ref[0], ('the oracle itself did not converge on this fixture', ref)
    assert o_end < 0.6 * ours[0] and o_end <= 2.0 * r_end + 0.3, (ours, ref)
    assert list(net.state_dict().keys()) == list(model.state_dict().keys())


def test_two_gpu_data_parallel_step(cuda):
    """NCCL gradient mean across 2 ranks (tools/ddp_check.py); skipped on a single-GPU box."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', str(29600 + os.getpid() % 300), os.path.join(root, 'tools', 'ddp_check.py')],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'DDP_CHECK ok' in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])


def test_neuston_net_train_then_run(cuda, tmp_path):
    """TRAIN entry point end to end on a small class-per-folder PNG tree (separable synthetic classes), then
    RUN with the checkpoint it wrote: files of the reference's layout appear and the model has learnt."""
    import json
    import os
    import numpy as np
    from PIL import Image
    from scipy.io import loadmat
    from oracle import synth_bins
    from tests.fixtures import class_rois
    from ifcb_classifier_b200 import neuston_net
    from ifcb_classifier_b200.neuston_models import NeustonModel
    src = tmp_path / 'dataset'
    imgs, labels = class_rois(96, 3, seed=5)
    for i, (im, k) in enumerate(zip(imgs, labels.tolist())):
        d = src / ('class_%d' % k)
        d.mkdir(parents=True, exist_ok=True)
        Image.fromarray(im, mode='L').save(str(d / ('img_%03d.png' % i)))
    out = tmp_path / 'train_out'
    # (batch 8 x 12 epochs = ~120 steps: BatchNorm running statistics need a few dozen steps at momentum 0.1
    #  before the eval-mode forward of the validation pass is meaningful -- same in the reference)
    rc = neuston_net.main(['--batch', '8', '--loaders', '2', 'TRAIN', str(src), 'resnet18', 'T1', '--untrain', '--seed', '7',
                           '--emax', '12', '--emin', '12', '--estop', '12', '--flip', 'xy', '--outdir', str(out)])
    assert rc == 0
    for f in ('T1.ptl', 'epochs.csv', 'args.yml', 'training_images.list', 'validation_images.list', 'results.mat'):
        assert (out / f).exists(), f
    rows = (out / 'epochs.csv').read_text().strip().splitlines()
    assert rows[0] == 'epoch,best,train_loss,val_loss,f1_macro,f1_weighted' and len(rows) >= 3
    n_train = len((out / 'training_images.list').read_text().splitlines())
    n_val = len((out / 'validation_images.list').read_text().splitlines())
    assert n_train + n_val == 96 and abs(n_val - 19) <= 3                      # --split 80:20 per class
    m = loadmat(str(out / 'results.mat'))
    assert m['confusion_matrix'].shape == (3, 3) and int(m['confusion_matrix'].sum()) == n_val
    assert m['output_scores'].shape == (n_val, 3) and m['training_classes'].min() >= 1     # list-typed index series are 1-based for MATLAB
    model = NeustonModel.load_from_checkpoint(str(out / 'T1.ptl'))
    assert model.hparams.classes == ['class_0', 'class_1', 'class_2'] and model.hparams.MODEL == 'resnet18'
    best = [r.split(',') for r in rows[1:] if r.split(',')[1] == 'True'][-1]
    assert float(best[5]) >= 0.6, rows                                          # weighted F1 of the best epoch: it learnt
    # RUN with the checkpoint
    bins = tmp_path / 'bins'
    synth_bins.write_bin(str(bins), synth_bins.make_bin(0, n_rois=20))
    rc = neuston_net.main(['--batch', '16', 'RUN', str(bins), str(out / 'T1.ptl'), 'R1', '--outdir', str(tmp_path / 'run_out'),
                           '--outfile', '{BIN_ID}_class.json'])
    assert rc == 0
    j = json.load(open(str(tmp_path / 'run_out' / (synth_bins.bin_lid(0) + '_class.json'))))
    assert len(j['output_classes']) == 20 and j['class_labels'] == ['class_0', 'class_1', 'class_2']
    # RUN --type img over the training tree: the model recognises its own classes
    rc = neuston_net.main(['--batch', '16', 'RUN', str(src), str(out / 'T1.ptl'), 'R2', '--type', 'img', '--outdir', str(tmp_path / 'run_img')])
    assert rc == 0
    j = json.load(open(str(tmp_path / 'run_img' / 'img_results.json')))
    assert len(j['input_images']) == 96 and 'bin_id' not in j
    truth = [int(os.path.basename(os.path.dirname(p_)).split('_')[1]) for p_ in j['input_images']]
    acc = float(np.mean(np.array(truth) == np.array(j['output_classes'])))
    assert acc >= 0.6, acc


def test_neuston_net_train_two_gpus(cuda, tmp_path):
    """`torchrun -m ifcb_classifier_b200.neuston_net ... TRAIN ...` on 2 GPUs: per-rank batches, NCCL gradient mean,
    rank 0 writes the reference's files; skipped on a single-GPU box."""
    import os, subprocess, sys
    from PIL import Image
    from tests.fixtures import class_rois
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    src = tmp_path / 'dataset'
    imgs, labels = class_rois(120, 3, seed=9)
    for i, (im, k) in enumerate(zip(imgs, labels.tolist())):
        d = src / ('class_%d' % k)
        d.mkdir(parents=True, exist_ok=True)
        Image.fromarray(im, mode='L').save(str(d / ('img_%03d.png' % i)))
    out = tmp_path / 'train_out'
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(29900 + os.getpid() % 90), '-m', 'ifcb_classifier_b200.neuston_net', '--batch', '8', '--loaders', '2',
           'TRAIN', str(src), 'resnet18', 'T2', '--untrain', '--seed', '5', '--emax', '6', '--emin', '6', '--estop', '6', '--outdir', str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    for f in ('T2.ptl', 'epochs.csv', 'args.yml', 'training_images.list', 'validation_images.list', 'results.mat'):
        assert (out / f).exists(), f
    rows = (out / 'epochs.csv').read_text().strip().splitlines()
    assert len(rows) == 7
    losses = [float(r_.split(',')[2]) for r_ in rows[1:]]
    assert losses[-1] < losses[0], rows


def test_bn_scratch_is_reusable_across_layers(cuda):
    """One 64 KB scratch serves every BN layer of a step: a narrow layer's backward (which leaves its coefficient
    floats in the scratch) must not disturb the accumulators a wider layer uses next, forward or backward."""
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.train import _vd
    L = _lib.lib()
    tdt, cdt = torch.bfloat16, _lib.IFCB_ACT_BF16
    g = torch.Generator().manual_seed(21)
    acc = torch.zeros(8192, dtype=torch.float64, device=cuda)

    def layer(Cc):
        z, z32 = _mk(cuda, 4, 6, 6, Cc, tdt, gen=g, scale=2.0)
        dy, dy32 = _mk(cuda, 4, 6, 6, Cc, tdt, gen=g)
        return dict(C=Cc, z=z, z32=z32, dy=dy, dy32=dy32, gamma=(torch.rand(Cc, generator=g) + 0.5).to(cuda),
                    beta=(torch.randn(Cc, generator=g) * 0.2).to(cuda), mean=torch.zeros(Cc, device=cuda), invstd=torch.zeros(Cc, device=cuda),
                    dgam=torch.zeros(Cc, device=cuda), dbet=torch.zeros(Cc, device=cuda))

    def stats(l):
        _lib.check(L.ifcb_bn_stats(C.byref(_vd(l['z'])), 4, cdt, 1e-5, 0.1, acc.data_ptr(), l['mean'].data_ptr(), l['invstd'].data_ptr(), None, None,
                                   _stream()), 'bn_stats')

    def backward(l):
        _lib.check(L.ifcb_bn_backward(C.byref(_vd(l['dy'])), None, C.byref(_vd(l['z'])), C.byref(_vd(l['dy'])), None, 0, 1, 4, cdt,
                                      l['mean'].data_ptr(), l['invstd'].data_ptr(), l['gamma'].data_ptr(), l['beta'].data_ptr(), acc.data_ptr(),
                                      l['dgam'].data_ptr(), l['dbet'].data_ptr(), _stream()), 'bn_backward')

    narrow, wide, wide2 = layer(64), layer(512), layer(2048)
    for l in (narrow, wide, wide2):
        stats(l)
    backward(narrow)                 # leaves 3*64 coefficient floats in the scratch
    stats(wide)                      # forward of a wider layer right after
    backward(wide)
    backward(wide2)
    torch.cuda.synchronize()
    for l in (wide, wide2):
        zr = l['z32'].clone().requires_grad_(True)
        gr, br = l['gamma'].clone().requires_grad_(True), l['beta'].clone().requires_grad_(True)
        y = F.relu(F.batch_norm(zr, None, None, gr, br, training=True, eps=1e-5))
        y.backward(l['dy32'])
        mean_ref = l['z32'].mean((0, 2, 3))
        assert torch.allclose(l['mean'], mean_ref, rtol=1e-4, atol=1e-5)
        assert torch.allclose(l['dbet'], br.grad, rtol=2e-3, atol=2e-3 * float(br.grad.abs().max()))
        assert torch.allclose(l['dgam'], gr.grad, rtol=2e-3, atol=2e-3 * float(gr.grad.abs().max()))
    assert float(acc[:4096].abs().max()) == 0.0


def test_cuda_graph_step_matches_eager(cuda):
    """The CUDA-graph replay of the step (train.TrainNet.enable_cuda_graph) runs the same kernels as the eager step:
    same losses over three Adam steps up to the order of the split-K gradient reductions."""
    from tests.fixtures import ref_model
    from ifcb_classifier_b200.train import TrainNet
    B, R = 16, 64
    model = ref_model('resnet18', 10, seed=6)
    g = torch.Generator().manual_seed(8)
    batches = [(torch.rand(B, 3, R, R, generator=g).to(cuda), torch.randint(0, 10, (B,), generator=g).to(cuda)) for _ in range(3)]
    losses, stats = {}, {}
    for mode in ('eager', 'graph'):
        net = TrainNet('resnet18', model.state_dict(), B, device=cuda, R=R, bucket_mb=4)
        if mode == 'graph':
            net.enable_cuda_graph()
            assert len(net._graph[0]) >= 3                        # one graph per gradient bucket
        losses[mode] = [float(net.step(x, y)) for x, y in batches]
        sd = net.state_dict()
        assert int(sd['bn1.num_batches_tracked']) == 3
        stats[mode] = sd['bn1.running_mean'], sd['layer3.0.bn1.running_var']
    assert abs(losses['eager'][0] - losses['graph'][0]) <= 1e-5 * abs(losses['eager'][0]), losses
    for a, b in zip(losses['eager'], losses['graph']):
        assert abs(a - b) <= 2e-2 * abs(a), losses
    for a, b in zip(stats['eager'], stats['graph']):              # capture warm-up leaves the running statistics alone
        assert torch.allclose(a, b, rtol=2e-2, atol=1e-3)


def test_stem_im2col_matches_unfold(cuda):
    """Patch matrix of the first conv (incl. torchvision's transform_input affine, inception.py:95-101): channel k =
    (r*kw + s)*3 + c of the 16-bit rounded, transformed input; zero outside the image and in the K padding."""
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.graph import transform_input_affine
    from ifcb_classifier_b200.train import _vd
    L = _lib.lib()
    g = torch.Generator().manual_seed(4)
    for (R, kh, st, pd, use_t) in ((37, 3, 2, 0, True), (30, 7, 2, 3, False)):
        x = torch.rand(3, 3, R, R, generator=g).to(cuda)
        P = (R + 2 * pd - kh) // st + 1
        K8 = (kh * kh * 3 + 7) // 8 * 8
        out, _ = _mk(cuda, 3, P, P, K8, torch.bfloat16, fill=False)
        ts, tb = transform_input_affine()
        sc, sh = ((C.c_float * 3)(*ts), (C.c_float * 3)(*tb)) if use_t else (None, None)
        _lib.check(L.ifcb_stem_im2col(x.data_ptr(), R, R, C.byref(_vd(out)), 3, kh, kh, st, pd, sc, sh, _lib.IFCB_ACT_BF16, _stream()), 'stem_im2col')
        xt = x
        if use_t:
            xt = torch.stack([torch.addcmul(torch.tensor(tb[c], device=cuda), x[:, c], torch.tensor(ts[c], device=cuda)) for c in range(3)], 1)
        # zero padding applies to the TRANSFORMED input (the module pads inside the conv)
        u = F.unfold(xt.to(torch.bfloat16).float(), (kh, kh), padding=pd, stride=st)
        u = u.view(3, 3, kh * kh, P, P).permute(0, 2, 1, 3, 4).reshape(3, kh * kh * 3, P, P)
        got = _read(out)
        assert float((got[:, :kh * kh * 3] - u).abs().max()) <= 2.0 ** -8 * float(u.abs().max())
        assert float(got[:, kh * kh * 3:].abs().max()) == 0.0


def test_short_last_batch_plan_shares_the_model(cuda):
    """The short last batch of an epoch is trained on AS IS (reference training_step, neuston_models.py:80-86) by a second
    plan of that size over the same parameter / gradient / Adam arenas: its loss matches the oracle's for the n samples (not
    for a wrapped-around full batch), it updates the shared parameters, and the main plan continues from them."""
    from oracle import train_ref
    from tests.fixtures import ref_model
    from ifcb_classifier_b200.train import TrainNet
    import copy
    B, n, R, C = 8, 5, 64, 6
    g = torch.Generator().manual_seed(11)
    model = ref_model('resnet18', C, seed=5).to(cuda)
    net = TrainNet('resnet18', model.state_dict(), B, device=cuda, R=R)
    tail = TrainNet('resnet18', model.state_dict(), n, device=cuda, R=R, share=net)
    assert tail.params.data_ptr() == net.params.data_ptr() and tail.m.data_ptr() == net.m.data_ptr()
    x8, y8 = torch.rand(B, 3, R, R, generator=g).to(cuda), torch.randint(0, C, (B,), generator=g).to(cuda)
    x5, y5 = torch.rand(n, 3, R, R, generator=g).to(cuda), torch.randint(0, C, (n,), generator=g).to(cuda)
    ref = copy.deepcopy(model)
    want = train_ref.train_steps(ref, [(x8, y8), (x5, y5), (x8, y8)])
    l0 = float(net.step(x8, y8))
    tail.repack()
    p_before = net.params.clone()
    l1 = float(tail.step(x5, y5))
    assert not torch.equal(p_before, net.params) and net.step_count == 2 and tail.step_count == 2
    net.repack()
    l2 = float(net.step(x8, y8))
    print('ours', [l0, l1, l2], 'oracle', want)
    assert abs(l0 - want[0]) <= 2e-2 * abs(want[0]) and abs(l1 - want[1]) <= 5e-2 * abs(want[1])
    sd, rsd = net.state_dict(), ref.state_dict()
    assert int(sd['bn1.num_batches_tracked']) == 3
    assert torch.allclose(sd['bn1.running_mean'], rsd['bn1.running_mean'].cpu(), rtol=5e-2, atol=5e-3)


def test_deterministic_mode_is_bitwise_reproducible(cuda):
    """Trainer(deterministic=True) upstream (neuston_net.py:101): with the deterministic workspace set, the weight gradient's
    split-K partials and the BatchNorm reductions are summed in a fixed order, so two runs of the same steps from the same
    state give bit-identical parameters (the default path combines them with floating-point atomics); and the deterministic
    result agrees with the default one to rounding."""
    from tests.fixtures import ref_model
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.train import TrainNet
    B, R, C = 16, 96, 7
    g = torch.Generator().manual_seed(21)
    model = ref_model('resnet50', C, seed=9)
    xs = [torch.rand(B, 3, R, R, generator=g).to(cuda) for _ in range(3)]
    ys = [torch.randint(0, C, (B,), generator=g).to(cuda) for _ in range(3)]

    def run(det):
        net = TrainNet('resnet50', model.state_dict(), B, device=cuda, R=R, deterministic=det, dropout=False)
        losses = [float(net.step(x, y)) for x, y in zip(xs, ys)]
        torch.cuda.synchronize()
        p, rm = net.params.clone(), net.buffers['layer3.2.bn2.running_var'].clone()
        del net
        return p, rm, losses

    try:
        a, ra, la = run(True)
        b, rb, lb = run(True)
        assert torch.equal(a, b) and torch.equal(ra, rb), 'deterministic mode is not reproducible'
    finally:
        _lib.check(_lib.lib().ifcb_train_deterministic(None, 0), 'train_deterministic off')
    c, rc_, lc = run(False)
    # same arithmetic up to the order of the split-K / block sums: three Adam steps stay close
    assert abs(la[0] - lc[0]) <= 1e-4 * abs(lc[0]) and torch.allclose(ra, rc_, rtol=5e-2, atol=1e-4)
    assert float((a - c).abs().max()) <= 1e-2            # (an Adam step moves a weight by ~lr whatever the gradient size: sign flips of ~0 gradients)


OTHER_FAMILIES = [('alexnet', 224, 8, 0.98), ('vgg11', 224, 4, 0.98), ('vgg11_bn', 224, 4, None), ('squeezenet', 96, 8, 0.97),
                  ('densenet121', 64, 8, None)]


@pytest.mark.parametrize('arch,R,B,min_cos', OTHER_FAMILIES, ids=[c[0] for c in OTHER_FAMILIES])
def test_train_step_other_families(cuda, arch, R, B, min_cos):
    """TRAIN for the remaining names get_namebrand_model resolves (reference neuston_models.py:27-42: alexnet, vgg*, squeezenet,
    densenet*): conv + bias + ReLU units without BatchNorm, Linear layers run as convolutions, Dropout, ceil-mode pools, DenseNet's
    pre-activation norms over the growing concatenation.  One training_step against the oracle (fp32 autograd of the same
    torchvision module, dropout off on both sides): loss within 2 %, every parameter has a gradient of the right shape, and for
    the BatchNorm-free families (no batch statistics to amplify 16-bit storage rounding) the whole gradient agrees with fp32
    (cosine); then a few Adam steps on the same batch bring the loss down, and state_dict() has torchvision's keys and shapes."""
    from oracle import train_ref
    from tests.fixtures import ref_model
    from ifcb_classifier_b200.train import TrainNet
    n_classes = 10
    model = ref_model(arch, n_classes, seed=3).to(cuda)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, R, R, generator=g).to(cuda)
    y = torch.randint(0, n_classes, (B,), generator=g).to(cuda)
    net = TrainNet(arch, model.state_dict(), B, device=cuda, dtype='bf16', dropout=False, R=R)
    loss = float(net.forward_backward(x, y))
    grads = {k: v.to(cuda).float() for k, v in net.grad_dict().items()}
    f_loss, f_grads = train_ref.forward_backward(model, x, y, dropout=False)
    assert sorted(grads) == sorted(f_grads)
    for k in f_grads:
        assert grads[k].shape == f_grads[k].shape, (k, grads[k].shape, f_grads[k].shape)
    dot = sum((grads[k] * f_grads[k]).sum() for k in f_grads)
    cos = float(dot / (torch.sqrt(sum((v ** 2).sum() for v in grads.values())) * torch.sqrt(sum((v ** 2).sum() for v in f_grads.values()))))
    print('%s: loss %.5f, fp32 oracle %.5f, gradient cosine %.4f' % (arch, loss, float(f_loss), cos))
    assert abs(loss - float(f_loss)) <= 2e-2 * abs(float(f_loss)), (loss, float(f_loss))
    if min_cos is not None:
        assert cos >= min_cos, cos
    sd, rsd = net.state_dict(), model.state_dict()
    assert list(sd.keys()) == list(rsd.keys())
    for k in rsd:
        assert tuple(sd[k].shape) == tuple(rsd[k].shape), k
    net.adam()
    losses = [float(net.step(x, y)) for _ in range(8)]
    assert min(losses[-3:]) < 0.9 * loss, (loss, losses)        # (Adam at lr 1e-3 is jumpy on the BatchNorm-free nets: a trend, not a rate)
    # with dropout on the step still runs and trains (mask from the counter-based generator)
    net2 = TrainNet(arch, model.state_dict(), B, device=cuda, dtype='bf16', dropout=True, R=R, seed=11)
    l2 = [float(net2.step(x, y)) for _ in range(3)]
    assert all(v == v and v < 50 for v in l2), l2
