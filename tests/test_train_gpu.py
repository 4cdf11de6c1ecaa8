"""TRAIN-step primitives against torch autograd (fp32, CPU/GPU) -- tolerances stated per test.
Everything goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

WGRAD_CASES = [
    # name, B, Cin, H, W, Cout, kh, kw, stride, pad
    ('3x3p1_64_64', 3, 64, 14, 14, 64, 3, 3, (1, 1), (1, 1)),
    ('1x1_256_64', 2, 256, 9, 9, 64, 1, 1, (1, 1), (0, 0)),
    ('3x3s2p1_64_128', 2, 64, 28, 28, 128, 3, 3, (2, 2), (1, 1)),
    ('1x1s2_64_256', 2, 64, 15, 15, 256, 1, 1, (2, 2), (0, 0)),
    ('3x3p1_128_192', 5, 128, 17, 17, 192, 3, 3, (1, 1), (1, 1)),
    ('1x7_160_160', 2, 160, 17, 17, 160, 1, 7, (1, 1), (0, 3)),
    ('7x7s2p3_8_64', 2, 8, 32, 32, 64, 7, 7, (2, 2), (3, 3)),
]


@pytest.mark.parametrize('dt', ['bf16', 'fp16'])
@pytest.mark.parametrize('case', WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_conv_wgrad(cuda, case, dt):
    from ifcb_classifier_b200 import _lib
    name, B, Cin, H, W, Cout, kh, kw, stride, pad = case
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
