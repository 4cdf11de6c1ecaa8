"""TRAIN step throughput on synthetic preprocessed tensors.

  python tools/bench_train.py --arch resnet50 --batch 256 --steps 10 --warmup 3
  torchrun --nproc-per-node 8 tools/bench_train.py --arch inception_v3 --batch 256

Per step: x float32 [B,3,R,R] already on the device (uniform random), labels uniform random
(torch.Generator().manual_seed(rank)), TrainNet.step = forward + backward + bucketed NCCL gradient
all-reduce (world > 1) + Adam + operand repack.  Timing: CUDA events on the launching stream, barrier
on both sides, max over ranks.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--arch', default='resnet50')
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--classes', type=int, default=100)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--parts', action='store_true', help='also time forward / backward / adam separately')
    ap.add_argument('--graph', action='store_true', help='replay the step from CUDA graphs')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    from ifcb_classifier_b200.neuston_models import get_namebrand_model
    from ifcb_classifier_b200.train import TrainNet
    torch.manual_seed(0)
    model = get_namebrand_model(args.arch, args.classes, pretrained=False)
    net = TrainNet(args.arch, model.state_dict(), args.batch, device=dev, dtype=args.dtype, seed=rank)
    g = torch.Generator().manual_seed(rank)
    x = torch.rand(args.batch, 3, net.R, net.R, generator=g).to(dev)
    y = torch.randint(0, args.classes, (args.batch,), generator=g).to(dev)
    net.inp.copy_(x)
    net.labels.copy_(y)
    if args.graph:
        net.enable_cuda_graph()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    losses = []
    for _ in range(args.warmup):
        losses.append(net.step().clone())
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses.append(net.step().clone())
    e1.record()
    sync()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    parts = {}
    if args.parts:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = [0.0, 0.0, 0.0]
        for _ in range(args.steps):
            ev[0].record(); net.forward(); ev[1].record(); net.backward(); ev[2].record(); net.adam(); ev[3].record()
            torch.cuda.synchronize()
            for i in range(3):
                acc[i] += ev[i].elapsed_time(ev[i + 1])
        parts = dict(forward_ms=acc[0] / args.steps, backward_ms=acc[1] / args.steps, adam_repack_ms=acc[2] / args.steps)
    if rank == 0:
        fwd_flops = net.fp.flops_per_image                       # 2*MACs of the train-mode forward (convs + stem + fc)
        img_s = args.batch * world / (ms / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('bf16_tflops_sustained', 1380.2))
        tf = 3 * fwd_flops * args.batch / (ms / 1e3) / 1e12      # per GPU
        print(json.dumps(dict(metric='TRAIN images/sec (%s %dpx, batch %d/GPU)' % (args.arch, net.R, args.batch), value=img_s, unit='img/s',
                              n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms, dtype=args.dtype, data='synthetic', cuda_graph=bool(args.graph),
                              gflop_per_img=3 * fwd_flops / 1e9, tflops_per_gpu=tf, frac_of_bf16_peak=tf / peak,
                              params=net.n_params, launches=dict(fwd=len(net.fwd), bwd=len(net.bwd)), loss_first=float(losses[0]),
                              loss_last=float(losses[-1]), mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, **parts)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
