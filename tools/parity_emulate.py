"""CPU emulation of operand-precision choices on the benchmarked configuration (Inception-v3, C = 100,
synthetic IFCB ROIs, fixture B = BN-calibrated random init): which tensors must keep more than 16 bits for the
top-1 agreement gate (>= 99.5 %) on the near-uniform softmax of fixture B?

Not a product path and not a test: a design experiment (torch fp32 CPU with rounding hooks).  It answers
VERDICT r01 "next" item 1(c) before GPU time is spent.  Results are recorded in DESIGN.md section 3.

  python tools/parity_emulate.py [--n 256] [--arch inception_v3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn


def rnd(t, fmt):
    if fmt == 'fp32':
        return t
    if fmt == 'fp16':
        return t.clamp(-65504, 65504).half().float()
    if fmt == 'bf16':
        return t.bfloat16().float()
    if fmt == 'fp16x2':                     # hi + lo split: value kept to ~22 bits
        hi = t.half().float()
        return hi + (t - hi).half().float()
    raise KeyError(fmt)


def emulate(model, x, w_fmt, a_fmt, precise_prefixes=(), precise_w='fp32', precise_a='fp32', batch=32):
    """Forward with conv weights rounded to w_fmt and every conv-bn-relu / pool output rounded to a_fmt; modules whose name
    starts with one of ``precise_prefixes`` use (precise_w, precise_a) instead."""
    import copy
    m = copy.deepcopy(model).eval()
    hooks = []
    for name, mod in m.named_modules():
        prec = any(name.startswith(p) for p in precise_prefixes)
        wf, af = (precise_w, precise_a) if prec else (w_fmt, a_fmt)
        if isinstance(mod, nn.Conv2d):
            with torch.no_grad():
                mod.weight.copy_(rnd(mod.weight, wf))
        # activations are stored 16-bit after BN+ReLU (BasicConv2d / resnet relu) and after pools
        if mod.__class__.__name__ == 'BasicConv2d' or isinstance(mod, (nn.MaxPool2d,)):
            hooks.append(mod.register_forward_hook(lambda _m, _i, o, af=af: rnd(o, af)))
    out = []
    with torch.no_grad():
        for i in range(0, x.shape[0], batch):
            out.append(torch.softmax(m(x[i:i + batch]), 1))
    for h in hooks:
        h.remove()
    return torch.cat(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=256)
    ap.add_argument('--arch', default='inception_v3')
    ap.add_argument('--classes', type=int, default=100)
    args = ap.parse_args()
    from oracle import synth_bins
    from oracle.pil_resize import ref_preprocess
    from tests import fixtures
    torch.set_num_threads(os.cpu_count() or 1)
    R = 299 if args.arch == 'inception_v3' else 224
    sb = synth_bins.make_bin(0, args.n)
    imgs = [sb['images'][t] for t in sorted(sb['images'])]
    x = torch.from_numpy(np.stack([ref_preprocess(im, R, None) for im in imgs]))
    model = fixtures.ref_model(args.arch, args.classes, seed=0)
    fixtures.calibrate_bn(model, x[:128], 'cpu')
    with torch.no_grad():
        ref = torch.cat([torch.softmax(model(x[i:i + 32]), 1) for i in range(0, args.n, 32)])
    top2 = ref.topk(2, 1).values
    margin = (top2[:, 0] - top2[:, 1])
    print('fixture B, %s C=%d, %d synthetic ROIs: mean max-prob %.4f, top1-top2 margin median %.2e, 5th pct %.2e, min %.2e'
          % (args.arch, args.classes, args.n, float(ref.max(1).values.mean()), float(margin.median()),
             float(margin.quantile(0.05)), float(margin.min())), flush=True)
    # fp32 noise floor on the CPU: the same fp32 graph with another summation order (channels_last kernels)
    with torch.no_grad():
        m2 = model.to(memory_format=torch.channels_last)
        alt = torch.cat([torch.softmax(m2(x[i:i + 32].contiguous(memory_format=torch.channels_last)), 1) for i in range(0, args.n, 32)])
    model = model.to(memory_format=torch.contiguous_format)

    def report(tag, got):
        print('%-60s top-1 agreement %.4f   max|dscore| %.3e' % (tag, float((ref.argmax(1) == got.argmax(1)).float().mean()),
                                                                  float((ref - got).abs().max())), flush=True)
    report('fp32 other summation order (CPU noise floor)', alt)
    last = ('Mixed_7b', 'Mixed_7c') if args.arch == 'inception_v3' else ('layer4',)
    late = ('Mixed_7a', 'Mixed_7b', 'Mixed_7c') if args.arch == 'inception_v3' else ('layer3', 'layer4')
    for tag, kw in [
        ('weights fp16, activations fp16 (round-1 default)', dict(w_fmt='fp16', a_fmt='fp16')),
        ('weights fp32, activations fp16', dict(w_fmt='fp32', a_fmt='fp16')),
        ('weights fp16, activations fp32', dict(w_fmt='fp16', a_fmt='fp32')),
        ('weights fp16x2 (hi+lo), activations fp16', dict(w_fmt='fp16x2', a_fmt='fp16')),
        ('weights fp16x2, activations fp16x2', dict(w_fmt='fp16x2', a_fmt='fp16x2')),
        ('fp16/fp16, last blocks %s fp32/fp32' % (last,), dict(w_fmt='fp16', a_fmt='fp16', precise_prefixes=last)),
        ('fp16/fp16, late blocks %s fp32/fp32' % (late,), dict(w_fmt='fp16', a_fmt='fp16', precise_prefixes=late)),
        ('weights bf16, activations bf16', dict(w_fmt='bf16', a_fmt='bf16')),
        ('weights fp32, activations bf16', dict(w_fmt='fp32', a_fmt='bf16')),
    ]:
        report(tag, emulate(model, x, **kw))


if __name__ == '__main__':
    main()
