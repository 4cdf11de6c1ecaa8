"""Restatement of the reference's TRAIN step (torch fp32 autograd).

TEST INFRASTRUCTURE (see oracle/__init__.py): only tests / smoke / bench's cpu_baseline import it.
Follows ``/root/reference/neuston_models.py``:
  * ``configure_optimizers`` (63-64): ``Adam(self.parameters(), lr=0.001)``.
  * ``training_step`` (81-86): train-mode forward of the torchvision model, then ``loss``.
  * ``loss`` (70-78): ``CE(logits, y) + 0.4 * CE(aux_logits, y)`` when the model returns the
    (logits, aux) pair (inception_v3 in train mode), plain ``CE`` otherwise.
and Lightning's automatic optimisation around it (zero_grad -> backward -> optimizer.step;
neuston_net.py:101-115).  The arithmetic is torch/torchvision (third-party, installed here).
"""
import torch
import torch.nn as nn

from .model_ref import loss as neuston_loss


def strict_fp32():
    """The oracle is fp32: switch off cuDNN / cuBLAS TF32 (10-bit significand) when it runs on a GPU."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def configure_optimizers(model):
    return torch.optim.Adam(model.parameters(), lr=0.001)


def training_step(model, x, y, dropout=True):
    """Train-mode forward + loss.  ``dropout=False`` turns the Inception head's Dropout(0.5) into
    the identity (its mask comes from torch's RNG stream, which no other implementation can
    reproduce) while BatchNorm stays in train mode."""
    model.train()
    if not dropout:
        for m in model.modules():
            if isinstance(m, nn.Dropout):
                m.eval()
    return neuston_loss(model(x), y)


def forward_backward(model, x, y, dropout=True):
    """One backward pass; returns (loss, {parameter name: gradient})."""
    strict_fp32()
    model.zero_grad(set_to_none=True)
    loss = training_step(model, x, y, dropout)
    loss.backward()
    return loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def train_steps(model, batches, dropout=True):
    """Adam steps over ``batches`` [(x, y), ...]; returns the list of losses (before each update)."""
    strict_fp32()
    opt = configure_optimizers(model)
    losses = []
    for x, y in batches:
        opt.zero_grad(set_to_none=True)
        loss = training_step(model, x, y, dropout)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return losses


# ------------------------------------------------------------------------------------------------
# Storage-rounding emulation.  The B200 path keeps activations and activation gradients in a
# 16-bit format between kernels (fp32 inside every kernel).  A ReLU network's gradient is not a
# continuous function of the forward values (a pre-activation that changes sign flips a mask), so
# comparing a 16-bit pipeline with pure fp32 autograd measures precision, not kernel correctness.
# ``with_storage_rounding`` inserts the SAME rounding points into the reference's own autograd
# graph (straight-through in both directions), which isolates the kernels' arithmetic.
# ------------------------------------------------------------------------------------------------
class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.dtype = dtype
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).to(g.dtype), None


def with_storage_rounding(model, dtype=torch.bfloat16):
    """Registers hooks on a torchvision resnet / inception_v3 so that conv outputs (z), block
    activations (a) and the gradients arriving at them are rounded to ``dtype`` exactly where the
    B200 path stores them; conv weights are rounded in place (master weights stay representable).
    Returns the list of hook handles."""
    import torchvision
    handles = []
    rnd = lambda mod, inp, out: _RoundSTE.apply(out, dtype)
    for name, m in model.named_modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data = m.weight.data.to(dtype).float()
            handles.append(m.register_forward_hook(rnd))                 # z
        elif isinstance(m, nn.ReLU):
            handles.append(m.register_forward_hook(rnd))                 # a (ResNet: relu module is shared)
        elif type(m).__name__ == 'BasicConv2d':
            handles.append(m.register_forward_hook(rnd))                 # a (Inception: F.relu inside)
        elif name.endswith('downsample'):
            handles.append(m.register_forward_hook(rnd))                 # identity branch activation (no ReLU)
    return handles
