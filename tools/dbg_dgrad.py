import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from ifcb_classifier_b200 import _lib
from ifcb_classifier_b200.graph import PlanBuilder, View
from ifcb_classifier_b200.train import build_dgrad
from tests.test_train_gpu import _mk, _read, _stream
cuda = torch.device('cuda:0')
name, B, Cin, H, W, Cout, kh, kw, stride, pad = ('5x5_128_768_to1x1', 4, 128, 5, 5, 768, 5, 5, (1, 1), (0, 0))
tdt = torch.bfloat16
g = torch.Generator().manual_seed(sum(name.encode()))
P = (H + 2 * pad[0] - kh) // stride[0] + 1
Q = (W + 2 * pad[1] - kw) // stride[1] + 1
w = (torch.randn(Cout, Cin, kh, kw, generator=g) * 0.05).to(tdt).float()
dy, dy32 = _mk(cuda, B, P, Q, Cout, tdt, gen=g)
dx, dx0 = _mk(cuda, B, H, W, Cin, tdt, gen=g)
bp = PlanBuilder(B, cuda, 'bf16')
dg = build_dgrad(bp, dy, dx, Cout, Cin, kh, kw, stride, pad, False)
master = w.permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin).contiguous().to(cuda)
_lib.check(_lib.lib().ifcb_conv_repack(master.data_ptr(), Cout, kh * kw, Cin, None, 0, dg['weight'].data_ptr(), dg['Cin_pad'], _lib.IFCB_ACT_BF16, _stream()), 'repack')
for r in dg['run']:
    r()
torch.cuda.synchronize()
xr = torch.zeros(B, Cin, H, W, device=cuda, requires_grad=True)
F.conv2d(xr, w.to(cuda), stride=stride, padding=pad).backward(dy32)
want = xr.grad
got = _read(dx)
# exact: double precision
want64 = torch.einsum('bo,oihw->bihw', dy32[:, :, 0, 0].double(), w.to(cuda).double())
err = (got - want).abs()
print('max err vs fp32 ref', float(err.max()), 'at', (err == err.max()).nonzero()[0].tolist())
e64 = (got.double() - want64).abs()
print('max err vs fp64', float(e64.max()), 'ref32 vs fp64', float((want.double() - want64).abs().max()))
i = (e64 == e64.max()).nonzero()[0].tolist()
print('ours', float(got[tuple(i)]), 'fp64', float(want64[tuple(i)]), 'fp32', float(want[tuple(i)]))
# error in units of bf16 half-ulp
ulp = 2.0 ** (torch.floor(torch.log2(want64.abs().clamp(min=1e-30))) - 7)
print('max err / ulp', float((e64 / ulp).max()), 'frac > 0.51 ulp', float(((e64 / ulp) > 0.51).float().mean()))
