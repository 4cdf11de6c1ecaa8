"""Layer-by-layer forward comparison TrainNet vs storage-rounded oracle (debugging aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from oracle import train_ref
from tests.fixtures import ref_model
from ifcb_classifier_b200.train import TrainNet

arch = sys.argv[1] if len(sys.argv) > 1 else 'resnet18'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
R = int(sys.argv[3]) if len(sys.argv) > 3 else 96
cuda = torch.device('cuda:0')
model = ref_model(arch, 10, seed=1).to(cuda)
train_ref.with_storage_rounding(model, torch.bfloat16)
zs, grads_z = {}, {}
for name, m in model.named_modules():
    if isinstance(m, nn.Conv2d):
        def hook(mod, inp, out, name=name):
            zs[name] = out.detach()
            out.register_hook(lambda g, name=name: grads_z.__setitem__(name, g.detach()))
        m.register_forward_hook(hook)
g = torch.Generator().manual_seed(7)
x = torch.rand(B, 3, R, R, generator=g).to(cuda)
y = torch.randint(0, 10, (B,), generator=g).to(cuda)
net = TrainNet(arch, model.state_dict(), B, device=cuda, dtype='bf16', dropout=False, R=R)
loss = float(net.forward_backward(x, y))
ref_loss, ref_grads = train_ref.forward_backward(model, x, y, dropout=False)
print('loss', loss, float(ref_loss))
for rec in net.records:
    if rec['kind'] != 'conv_bn':
        continue
    nm = rec['name']
    ours = rec['z'].interior().float().permute(0, 3, 1, 2)
    ref = zs[nm]
    d = (ours - ref)
    # dz: after backward our grad_of(out) holds dz (in place)
    dz = net.grad_of(rec['out']).interior().float().permute(0, 3, 1, 2)
    gz = grads_z[nm]
    print('%-34s z rel %.5f maxabs %.4f (|z| %.3f) frac>1ulp %.5f | dz rel %.5f cos %.5f' % (
        nm, float(d.norm() / ref.norm()), float(d.abs().max()), float(ref.abs().max()),
        float((d.abs() > 2.0 ** -7 * ref.abs().clamp(min=1e-3)).float().mean()),
        float((dz - gz).norm() / gz.norm()), float((dz * gz).sum() / (dz.norm() * gz.norm()))))
