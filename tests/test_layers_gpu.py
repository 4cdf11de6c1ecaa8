"""Layer-level parity of the sm_100a kernels against plain PyTorch fp32 (CPU) --
floating-point kernels keep a torch reference; tolerances are stated per test.
Everything goes through the C ABI via ifcb_classifier_b200.graph.PlanBuilder."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf16r(x):
    return x.bfloat16().float()


def nhwc_dev(x_nchw, dev, c_total=None, c0=0):
    """NCHW fp32 -> NHWC bf16 cuda tensor (optionally embedded at channel c0 of a wider tensor)."""
    B, Cc, H, W = x_nchw.shape
    c_total = c_total or Cc
    t = torch.full((B, H, W, c_total), 7.0, dtype=torch.bfloat16, device=dev)
    t[..., c0:c0 + Cc] = x_nchw.permute(0, 2, 3, 1).to(dev).bfloat16()
    return t


def to_nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2)


def padded_view(x_nchw, dev, pad, cap=None, dtype=torch.bfloat16):
    """NCHW fp32 -> View over a zero-bordered NHWC device tensor (batch capacity `cap`)."""
    from ifcb_classifier_b200.graph import View
    B, Cc, H, W = x_nchw.shape
    cap = cap or B
    t = torch.zeros((cap, H + 2 * pad[0], W + 2 * pad[1], Cc), dtype=dtype, device=dev)
    t[:B, pad[0]:pad[0] + H, pad[1]:pad[1] + W] = x_nchw.permute(0, 2, 3, 1).to(dev).to(dtype)
    return View(t, pad=pad)


def border_is_zero(v):
    t = v.t.float().clone()
    ph, pw = v.pad
    t[:, ph:t.shape[1] - ph, pw:t.shape[2] - pw] = 0
    return float(t.abs().max()) == 0.0


def swizzle_decode(raw_u8, row_bytes=128):
    """raw swizzled operand tile (128 rows of one swizzle span: 128 B = SWIZZLE_128B, 64 B =
    SWIZZLE_64B) -> [128 rows, row_bytes/2] bf16 values (as float).  16-byte chunk j of the row
    at byte address a sits at chunk j ^ ((a >> 7) & (chunks - 1))."""
    chunks = row_bytes // 16
    raw = raw_u8[:128 * row_bytes].view(128, chunks, 16)
    rows = torch.arange(128)
    phase = ((rows * row_bytes) >> 7) & (chunks - 1)
    out = torch.empty_like(raw)
    for j in range(chunks):
        out[rows, j] = raw[rows, (j ^ phase)]
    return out.reshape(128, row_bytes).contiguous().view(torch.bfloat16).float()


CONV_CASES = [
    # name, B, Cin, H, W, Cout, kh, kw, stride, pad
    ('1x1_64_80', 3, 64, 13, 13, 80, 1, 1, (1, 1), (0, 0)),
    ('3x3p1_64_96', 2, 64, 17, 17, 96, 3, 3, (1, 1), (1, 1)),
    ('3x3s2_288_384', 2, 288, 35, 35, 384, 3, 3, (2, 2), (0, 0)),
    ('5x5p2_48_64', 1, 48, 35, 35, 64, 5, 5, (1, 1), (2, 2)),
    ('1x7_128_128', 3, 128, 17, 17, 128, 1, 7, (1, 1), (0, 3)),
    ('7x1_160_192', 3, 160, 17, 17, 192, 7, 1, (1, 1), (3, 0)),
    ('3x3p1_448_384', 5, 448, 8, 8, 384, 3, 3, (1, 1), (1, 1)),
    ('3x3_32_32', 2, 32, 21, 21, 32, 3, 3, (1, 1), (0, 0)),
    ('3x3p1_32_64', 2, 32, 19, 19, 64, 3, 3, (1, 1), (1, 1)),
    ('3x3s2_16_48', 2, 16, 19, 19, 48, 3, 3, (2, 2), (0, 0)),
    ('3x3_80_192', 1, 80, 23, 23, 192, 3, 3, (1, 1), (0, 0)),
    ('1x1s2_64_128', 2, 64, 56, 56, 128, 1, 1, (2, 2), (0, 0)),
    ('3x3s2p1_64_128', 2, 64, 28, 28, 128, 3, 3, (2, 2), (1, 1)),
    ('1x3_384_384', 2, 384, 8, 8, 384, 1, 3, (1, 1), (0, 1)),
    ('1x1_2048_320', 9, 2048, 8, 8, 320, 1, 1, (1, 1), (0, 0)),
]


def _ref_conv(x, w, scale, shift, stride, pad, relu, residual=None):
    y = F.conv2d(bf16r(x), bf16r(w), stride=stride, padding=pad)
    y = y * scale[None, :, None, None] + shift[None, :, None, None]
    if residual is not None:
        y = y + bf16r(residual)
    return torch.relu(y) if relu else y


def _check(got, want, what):
    # bf16 output rounding (2^-9 relative) + fp32 accumulation-order noise
    err = (got - want).abs()
    tol = 1.5e-2 * want.abs() + 2e-2
    assert bool((err <= tol).all()), '%s: max err %.4g at want %.4g' % (
        what, float(err.max()), float(want.flatten()[err.argmax()]))


def test_im2col_tma_probe(cuda):
    """The im2col TMA load delivers pixel-major rows of 64 channels with zero fill for padding."""
    import ctypes as C
    from ifcb_classifier_b200 import _lib
    from ifcb_classifier_b200.graph import PlanBuilder, View
    torch.manual_seed(0)
    for (Cin, H, W, kh, kw, stride, pad) in [(64, 10, 10, 3, 3, (1, 1), (1, 1)), (32, 9, 11, 3, 3, (2, 2), (0, 0)),
                                             (80, 7, 7, 1, 7, (1, 1), (0, 3))]:
        B = 4
        x = torch.randn(B, Cin, H, W)
        pb = PlanBuilder(B, cuda, 'bf16')
        xin = View(nhwc_dev(x, cuda)); pb.keep.append(xin.t)
        P = (H + 2 * pad[0] - kh) // stride[0] + 1
        Q = (W + 2 * pad[1] - kw) // stride[1] + 1
        out = pb.alloc(P, Q, 16)
        pb.conv(xin, [dict(weight=torch.randn(16, Cin, kh, kw), scale=torch.ones(16), shift=torch.zeros(16),
                           relu=False, out=out)], stride, pad, algo=1)
        raw = torch.zeros(16384, dtype=torch.uint8, device=cuda)
        xp = F.pad(bf16r(x), (pad[1], pad[1], pad[0], pad[0]))
        for (m0, r, s, cb) in [(0, 0, 0, 0), (128, kh - 1, kw - 1, 0), (P * Q - 5, 0, kw - 1, 0 if Cin <= 32 else (Cin - 1) // 64)]:
            img, rem = divmod(m0, P * Q)
            op, oq = divmod(rem, Q)
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().ifcb_debug_im2col_probe(pb.handle, 0, cb * 64, oq * stride[1] - pad[1],
                                                          op * stride[0] - pad[0], img, s, r, raw.data_ptr(), stream))
            torch.cuda.synchronize()
            relems = 32 if Cin <= 32 else 64          # thin layers use 64-byte operand rows
            tile = swizzle_decode(raw.cpu(), 2 * relems)
            want = torch.zeros(128, relems)
            for i in range(128):
                m = m0 + i
                n_, rem_ = divmod(m, P * Q)
                p_, q_ = divmod(rem_, Q)
                if n_ >= B:
                    continue
                c1 = min(Cin, cb * 64 + 64)
                want[i, :c1 - cb * 64] = xp[n_, cb * 64:c1, p_ * stride[0] + r, q_ * stride[1] + s]
            assert torch.equal(tile, want), (Cin, H, W, kh, kw, m0, r, s, cb, float((tile - want).abs().max()))
        pb.close()


@pytest.mark.parametrize('algo', ['im2col', 'window', 'pair'])
@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_bn_relu(cuda, case, algo):
    from ifcb_classifier_b200.graph import PlanBuilder
    from ifcb_classifier_b200._lib import IFCB_CONV_IM2COL, IFCB_CONV_WINDOW, IFCB_CONV_IM2COL_PAIR
    name, B, Cin, H, W, Cout, kh, kw, stride, pad = case
    if algo == 'window' and stride != (1, 1):
        pytest.skip('window algorithm is stride-1 only')
    g = torch.Generator().manual_seed(sum(name.encode()))
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, kh, kw, generator=g) / np.sqrt(Cin * kh * kw)
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.2
    cap = B + 1                                            # run with batch < capacity
    pb = PlanBuilder(cap, cuda, 'bf16')
    xin = padded_view(x, cuda, pad if algo == 'window' else (0, 0), cap)
    pb.keep.append(xin.t)
    P = (H + 2 * pad[0] - kh) // stride[0] + 1
    Q = (W + 2 * pad[1] - kw) // stride[1] + 1
    out = pb.alloc(P, Q, Cout, pad=(1, 2) if algo != 'im2col' else (0, 0))     # padded destination
    pb.conv(xin, [dict(weight=w, scale=scale, shift=shift, relu=True, out=out)], stride, pad, name=name,
            algo={'window': IFCB_CONV_WINDOW, 'im2col': IFCB_CONV_IM2COL, 'pair': IFCB_CONV_IM2COL_PAIR}[algo])
    pb.run(B)
    torch.cuda.synchronize()
    got = to_nchw(out.interior()[:B])
    want = _ref_conv(x, w, scale, shift, stride, pad, True)
    _check(got, want, name)
    assert float(out.t[B:].float().abs().max()) == 0.0     # rows past the batch are untouched
    assert border_is_zero(out)                             # junk rows never reach the zero border
    pb.close()


def test_window_input_padded_more_than_conv(cuda):
    """Inception-E pattern: one (1,1)-padded tensor feeds a 1x3 (pad 0,1) and a 3x1 (pad 1,0) conv."""
    from ifcb_classifier_b200.graph import PlanBuilder
    from ifcb_classifier_b200._lib import IFCB_CONV_WINDOW
    g = torch.Generator().manual_seed(31)
    B, Cin, H, W, Cout = 5, 384, 8, 8, 384
    x = torch.randn(B, Cin, H, W, generator=g)
    pb = PlanBuilder(B, cuda, 'bf16')
    xin = padded_view(x, cuda, (1, 1)); pb.keep.append(xin.t)
    cat = pb.alloc(H, W, 2 * Cout)
    ws = []
    for i, (k, pad) in enumerate((((1, 3), (0, 1)), ((3, 1), (1, 0)))):
        w = torch.randn(Cout, Cin, k[0], k[1], generator=g) / np.sqrt(Cin * 3)
        ws.append((w, pad))
        pb.conv(xin, [dict(weight=w, scale=torch.ones(Cout), shift=torch.zeros(Cout), relu=True,
                           out=cat.slice(i * Cout, (i + 1) * Cout))], (1, 1), pad, algo=IFCB_CONV_WINDOW)
    pb.run(B)
    torch.cuda.synchronize()
    for i, (w, pad) in enumerate(ws):
        want = _ref_conv(x, w, torch.ones(Cout), torch.zeros(Cout), (1, 1), pad, True)
        _check(to_nchw(cat.t[..., i * Cout:(i + 1) * Cout]), want, 'E-branch %d' % i)
    pb.close()


def test_conv_fused_segments_slices_and_residual(cuda):
    """Horizontally fused 1x1s scattering to channel slices (one raw segment), input read
    from a channel slice of a wider tensor, and a residual add (ResNet block tail)."""
    from ifcb_classifier_b200.graph import PlanBuilder, View
    g = torch.Generator().manual_seed(5)
    B, H, W, Cin = 3, 35, 35, 192
    x = torch.randn(B, Cin, H, W, generator=g)
    couts = [64, 48, 64, 32]
    ws = [torch.randn(c, Cin, 1, 1, generator=g) / np.sqrt(Cin) for c in couts]
    scs = [torch.rand(c, generator=g) + 0.5 for c in couts]
    shs = [torch.randn(c, generator=g) * 0.2 for c in couts]
    pb = PlanBuilder(B, cuda, 'bf16')
    wide = nhwc_dev(x, cuda, c_total=256, c0=32); pb.keep.append(wide)
    xin = View(wide, 32, 32 + Cin)
    cat = pb.alloc(H, W, 256)
    outs = [cat.slice(0, 64), pb.alloc(H, W, 48, pad=(2, 2)), cat.slice(128, 192), pb.alloc(H, W, 32, pad=(0, 3))]
    relus = [True, True, True, False]
    pb.conv(xin, [dict(weight=ws[i], scale=scs[i], shift=shs[i], relu=relus[i], out=outs[i]) for i in range(4)])
    # residual: 1x1 64 -> 256 on top of `res`
    xr = torch.randn(B, 64, 14, 14, generator=g)
    res = torch.randn(B, 256, 14, 14, generator=g)
    wr = torch.randn(256, 64, 1, 1, generator=g) / 8
    xr_d, res_d = View(nhwc_dev(xr, cuda)), padded_view(res, cuda, (1, 1))
    pb.keep += [xr_d.t, res_d.t]
    out_r = pb.alloc(14, 14, 256)
    pb.conv(xr_d, [dict(weight=wr, scale=torch.ones(256), shift=torch.zeros(256), relu=True, out=out_r)],
            residual=res_d)
    pb.run(B)
    torch.cuda.synchronize()
    for i in range(4):
        want = _ref_conv(x, ws[i], scs[i], shs[i], (1, 1), (0, 0), relus[i])
        o = outs[i]
        got = to_nchw(o.interior())
        _check(got, want, 'segment %d' % i)
        assert border_is_zero(o) or o.t is cat.t
    assert float(cat.t[..., 64:128].float().abs().max()) == 0.0     # untouched slice of the concat buffer
    want = _ref_conv(xr, wr, torch.ones(256), torch.zeros(256), (1, 1), (0, 0), True, residual=res)
    _check(to_nchw(out_r.t), want, 'residual')
    pb.close()


@pytest.mark.parametrize('algo', ['im2col', 'window', 'pair'])
def test_large_batch_many_tiles(cuda, algo):
    """More tiles than SMs: exercises the persistent loop, both TMEM buffers and phase wrap."""
    from ifcb_classifier_b200.graph import PlanBuilder
    g = torch.Generator().manual_seed(11)
    B, H, W, Cin, Cout = 40, 35, 35, 96, 96
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / np.sqrt(Cin * 9)
    pb = PlanBuilder(B, cuda, 'bf16')
    xin = padded_view(x, cuda, (1, 1) if algo == 'window' else (0, 0)); pb.keep.append(xin.t)
    out = pb.alloc(H, W, Cout)
    pb.conv(xin, [dict(weight=w, scale=torch.ones(Cout), shift=torch.zeros(Cout), relu=False, out=out)],
            (1, 1), (1, 1), algo={'window': 2, 'im2col': 1, 'pair': 3}[algo])
    pb.run(B)
    torch.cuda.synchronize()
    _check(to_nchw(out.t), _ref_conv(x, w, torch.ones(Cout), torch.zeros(Cout), (1, 1), (1, 1), False), 'many tiles')
    pb.close()


def test_pools(cuda):
    from ifcb_classifier_b200.graph import PlanBuilder, View
    from ifcb_classifier_b200._lib import IFCB_POOL_MAX, IFCB_POOL_AVG_AFFINE
    g = torch.Generator().manual_seed(2)
    B = 3
    x = torch.randn(B, 64, 37, 37, generator=g)
    pb = PlanBuilder(B, cuda, 'bf16')
    xin = View(nhwc_dev(x, cuda)); pb.keep.append(xin.t)
    o1 = pb.alloc(18, 18, 64)
    pb.pool(IFCB_POOL_MAX, xin, 3, 2, 0, o1)
    o2 = pb.alloc(19, 19, 64, pad=(1, 1))
    pb.pool(IFCB_POOL_MAX, xin, 3, 2, 1, o2)
    xin_p = padded_view(x, cuda, (2, 1)); pb.keep.append(xin_p.t)
    o3 = pb.alloc(18, 18, 64)
    pb.pool(IFCB_POOL_MAX, xin_p, 3, 2, 0, o3)
    cat = pb.alloc(37, 37, 96)
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.3
    pb.pool(IFCB_POOL_AVG_AFFINE, xin, 3, 1, 1, cat.slice(16, 80), sc, sh, relu=True)
    pb.run(B)
    torch.cuda.synchronize()
    xb = bf16r(x)
    assert torch.equal(to_nchw(o1.t), F.max_pool2d(xb, 3, 2, 0))
    assert torch.equal(to_nchw(o2.interior()), F.max_pool2d(xb, 3, 2, 1)) and border_is_zero(o2)
    assert torch.equal(to_nchw(o3.t), F.max_pool2d(xb, 3, 2, 0))
    want = torch.relu(F.avg_pool2d(xb, 3, 1, 1) * sc[None, :, None, None] + sh[None, :, None, None])
    _check(to_nchw(cat.t[..., 16:80]), want, 'avgpool affine')
    pb.close()


def test_stem_u8_and_f32(cuda):
    from ifcb_classifier_b200.graph import PlanBuilder, input_lut, input_affine
    from ifcb_classifier_b200._lib import IFCB_STEM_IN_U8_GRAY, IFCB_STEM_IN_F32_NCHW
    g = torch.Generator().manual_seed(3)
    B, R = 2, 61
    gray = torch.randint(0, 256, (B, R, R), generator=g, dtype=torch.uint8)
    mean, std = [0.5, 0.4, 0.3], [0.2, 0.25, 0.3]
    lut = input_lut((mean, std), False)
    x = torch.stack([lut[c][gray.long()] for c in range(3)], 1)              # [B,3,R,R] float32
    # (32, 3, s, 0) from u8 is the tensor-core stem (Inception Conv2d_1a; 1800 / 6962 pixels: partial last tile)
    for (Co, k, stride, pad) in [(32, 3, 2, 0), (32, 3, 1, 0), (64, 7, 2, 3)]:
        w = torch.randn(Co, 3, k, k, generator=g) / np.sqrt(3 * k * k)
        sc, sh = torch.rand(Co, generator=g) + 0.5, torch.randn(Co, generator=g) * 0.2
        want = torch.relu(F.conv2d(x, w, stride=stride, padding=pad) * sc[None, :, None, None] + sh[None, :, None, None])
        P = (R + 2 * pad - k) // stride + 1
        for kind, inp in ((IFCB_STEM_IN_U8_GRAY, gray.to(cuda)), (IFCB_STEM_IN_F32_NCHW, x.to(cuda))):
            for dt in ('bf16', 'fp16'):
                pb = PlanBuilder(B, cuda, dt)
                pb.keep.append(inp)
                out = pb.alloc(P, P, Co)
                pb.stem(inp, kind, R, R, w, sc, sh, stride, pad, out, affine=input_affine((mean, std), False))
                pb.run(B)
                torch.cuda.synchronize()
                _check(to_nchw(out.t), want, 'stem %d stride %d kind %d %s' % (Co, stride, kind, dt))
                if kind == IFCB_STEM_IN_U8_GRAY and k == 3:
                    # the accumulation is fp32-exact (three-term weight split): only the 16-bit output rounding remains
                    err = (to_nchw(out.t) - want).abs()
                    assert bool((err <= 2.0 ** (-8 if dt == 'bf16' else -11) * want.abs() + 1e-5).all()), (stride, dt, float(err.max()))
                pb.close()


def test_head_softmax_top1(cuda):
    from ifcb_classifier_b200.graph import PlanBuilder, View
    g = torch.Generator().manual_seed(4)
    B, Cc, HW, K = 5, 2048, 64, 100
    x = torch.randn(B, Cc, 8, 8, generator=g)
    w = torch.randn(K, Cc, generator=g) / np.sqrt(Cc)
    b = torch.randn(K, generator=g) * 0.1
    pb = PlanBuilder(B, cuda, 'bf16')
    xin = View(nhwc_dev(x, cuda)); pb.keep.append(xin.t)
    pb.head(xin, w, b)
    pb.run(B)
    torch.cuda.synchronize()
    pooled = bf16r(x).mean((2, 3))
    logits = pooled @ w.t() + b
    want = torch.softmax(logits, 1)
    got = pb.scores.cpu()
    assert float((got - want).abs().max()) < 1e-5          # fp32 head: summation order only
    assert float((pb.logits.cpu() - logits).abs().max()) < 1e-4
    assert torch.equal(pb.top1.cpu().long(), got.argmax(1))
    assert torch.equal(pb.top1_score.cpu(), got.max(1).values)
    assert float((got.sum(1) - 1).abs().max()) < 1e-5
    pb.close()


def test_conv_fp16_operands(cuda):
    """Same kernel with fp16 operands/activations (the default RUN format)."""
    from ifcb_classifier_b200.graph import PlanBuilder, View
    g = torch.Generator().manual_seed(21)
    B, Cin, H, W, Cout = 3, 160, 17, 17, 192
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 1, 7, generator=g) / np.sqrt(Cin * 7)
    sc, sh = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.2
    pb = PlanBuilder(B, cuda, 'fp16')
    xin = View(x.permute(0, 2, 3, 1).contiguous().to(cuda).half()); pb.keep.append(xin.t)
    out = pb.alloc(H, W, Cout)
    pb.conv(xin, [dict(weight=w, scale=sc, shift=sh, relu=True, out=out)], (1, 1), (0, 3))
    pb.run(B)
    torch.cuda.synchronize()
    y = F.conv2d(x.half().float(), w.half().float(), padding=(0, 3))
    want = torch.relu(y * sc[None, :, None, None] + sh[None, :, None, None])
    err = (to_nchw(out.t) - want).abs()
    assert bool((err <= 2e-3 * want.abs() + 2e-3).all()), float(err.max())      # fp16: 2^-11 relative
    pb.close()
