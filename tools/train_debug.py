"""Per-parameter gradient comparison of TrainNet vs the oracle (debugging aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import train_ref
from tests.fixtures import ref_model
from ifcb_classifier_b200.train import TrainNet

arch = sys.argv[1] if len(sys.argv) > 1 else 'resnet18'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
R = int(sys.argv[3]) if len(sys.argv) > 3 else 96
dt = sys.argv[4] if len(sys.argv) > 4 else 'bf16'
cuda = torch.device('cuda:0')
model = ref_model(arch, 10, seed=1).to(cuda)
if os.environ.get('QREF', '1') == '1':
    train_ref.with_storage_rounding(model, torch.bfloat16 if dt == 'bf16' else torch.float16)
g = torch.Generator().manual_seed(7)
x = torch.rand(B, 3, R, R, generator=g).to(cuda)
y = torch.randint(0, 10, (B,), generator=g).to(cuda)
net = TrainNet(arch, model.state_dict(), B, device=cuda, dtype=dt, dropout=False, R=R)
loss = float(net.forward_backward(x, y))
grads = net.grad_dict()
ref_loss, ref_grads = train_ref.forward_backward(model, x, y, dropout=False)
print('loss', loss, float(ref_loss))
for k in reversed(list(ref_grads)):
    a, b = grads[k].to(cuda).float(), ref_grads[k]
    print('%-40s cos %.4f rel %.4f |ref| %.3e' % (k, float((a * b).sum() / (a.norm() * b.norm() + 1e-30)),
                                                   float((a - b).norm() / (b.norm() + 1e-30)), float(b.norm())))
