"""Data-parallel TRAIN check on >= 2 GPUs (torchrun --nproc-per-node 2 tools/ddp_check.py).

Reference semantics (Lightning DDP, neuston_net.py:101-107): every rank steps on its own batch with its
own BatchNorm statistics, gradients are AVERAGED over ranks, every rank applies the same Adam update.
Checks: (1) bucketed all-reduce result == mean of the ranks' local gradients, (2) parameters stay
bit-identical across ranks through overlapped steps, (3) the loss goes down.  Prints 'DDP_CHECK ok'."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from ifcb_classifier_b200.neuston_models import get_namebrand_model
    from ifcb_classifier_b200.sharding import GradReducer
    from ifcb_classifier_b200.train import TrainNet
    torch.manual_seed(0)
    model = get_namebrand_model('resnet18', 10, pretrained=False)
    B, R = 16, 64
    net = TrainNet('resnet18', model.state_dict(), B, device=dev, dtype='bf16', R=R, bucket_mb=4)
    g = torch.Generator().manual_seed(1234 + rank)
    x = torch.rand(B, 3, R, R, generator=g).to(dev)
    y = torch.randint(0, 10, (B,), generator=g).to(dev)
    assert len(net.bucket_marks) >= 3, net.bucket_marks
    # (1) local gradients, then the exchange
    net.forward_backward(x, y)
    local_g = net.grads.clone()
    gathered = [torch.empty_like(local_g) for _ in range(world)]
    dist.all_gather(gathered, local_g)
    want = sum(gathered) / world
    red = GradReducer(net.grads)
    for _, lo, hi in net.bucket_marks:
        red(lo, hi)
    scale = red.wait()
    torch.cuda.synchronize()
    err = float((net.grads * scale - want).abs().max())
    assert err <= 1e-6 * float(want.abs().max()) + 1e-12, err
    net.adam(scale)
    # (2) overlapped steps keep the replicas identical
    losses = [float(net.step(x, y)) for _ in range(4)]
    params = [torch.empty_like(net.params) for _ in range(world)]
    dist.all_gather(params, net.params)
    for p in params[1:]:
        assert torch.equal(p, params[0]), 'replicas diverged'
    assert losses[-1] < losses[0], losses
    if rank == 0:
        print('DDP_CHECK ok world=%d buckets=%d allreduce_err=%.2e losses=%s' % (world, len(net.bucket_marks), err, losses))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
