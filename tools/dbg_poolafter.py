import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from tests.fixtures import ref_model
from ifcb_classifier_b200.train import TrainNet
from tests.train_local import _nchw
cuda = torch.device('cuda:0')
model = ref_model('inception_v3', 10, seed=1).to(cuda)
g = torch.Generator().manual_seed(7)
B = 8
x = torch.rand(B, 3, 299, 299, generator=g).to(cuda)
y = torch.randint(0, 10, (B,), generator=g).to(cuda)
net = TrainNet('inception_v3', model.state_dict(), B, device=cuda, dtype='bf16', dropout=False, keep_dy=True)
net.forward_backward(x, y)
torch.cuda.synchronize()
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
for rec in net.records:
    if rec['kind'] == 'conv_bn' and rec.get('pool_after') is not None:
        xin = _nchw(rec['x'])
        dz = _nchw(rec['dz'])
        dr = _nchw(rec['dr'])
        t = torch.zeros_like(dr, requires_grad=True)
        F.avg_pool2d(t, 3, 1, 1).backward(dz)
        dr_exp = t.grad
        Co, Ci = rec['Co'], rec['Ci']
        dW = rec['pw'].g[:, 0, :Ci]
        dW_dr = torch.einsum('bohw,bihw->oi', dr, xin)
        dW_dz = torch.einsum('bohw,bihw->oi', dz, xin)
        dW_exp = torch.einsum('bohw,bihw->oi', dr_exp, xin)
        print('%-28s dr rel %.4f | dW vs x*dr %.4f, vs x*dz(unpooled) %.4f, vs expected %.4f | |dW| %.3e' % (
            rec['name'], rel(dr, dr_exp), rel(dW, dW_dr), rel(dW, dW_dz), rel(dW, dW_exp), float(dW.norm())))
print('---- autograd expectation as tests/train_local.py builds it')
q = lambda t: t.to(torch.bfloat16).float()
for rec in net.records:
    if rec['kind'] == 'conv_bn' and rec.get('pool_after') is not None:
        Co, Ci, kh, kw = rec['Co'], rec['Ci'], rec['kh'], rec['kw']
        pw = rec['pw']
        w = pw.w[:, :, :Ci].reshape(Co, kh, kw, Ci).permute(0, 3, 1, 2).contiguous()
        x_in, w_op = _nchw(rec['x']), q(w)
        x_in.requires_grad_(True); w_op.requires_grad_(True)
        z_exp = F.conv2d(x_in, w_op, stride=rec['stride'], padding=rec['pad'])
        z_exp = F.avg_pool2d(z_exp, *rec['pool_after'])
        dz = _nchw(rec['dz'])
        z_exp.backward(dz)
        dW = pw.g[:, :, :Ci].reshape(Co, kh, kw, Ci).permute(0, 3, 1, 2)
        t = torch.zeros_like(dz, requires_grad=True)
        F.avg_pool2d(t, 3, 1, 1).backward(dz)
        dW_exp = torch.einsum('bohw,bihw->oi', t.grad, x_in.detach())
        print('%-28s ours vs autograd %.4f, autograd vs einsum %.4f, shapes %s %s stride %s pad %s pool %s' % (
            rec['name'], rel(dW, w_op.grad), rel(w_op.grad[:, :, 0, 0], dW_exp), tuple(dW.shape), tuple(w_op.grad.shape), rec['stride'], rec['pad'], rec['pool_after']))
