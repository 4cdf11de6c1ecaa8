"""Single-layer timing matrix for the conv kernel (profiling aid, not a bench line).
Times representative Inception-v3 layers at batch 512 under both algorithms and the
debug switches of IFCB_CONV_DEBUG (1 = skip TMA loads, 2 = skip MMAs, 4 = skip epilogue)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifcb_classifier_b200.graph import PlanBuilder, View  # noqa: E402

dev = torch.device('cuda:0')
B = int(os.environ.get('BL_BATCH', '512'))
LAYERS = [
    ('2a 3x3 32->32 @149', 32, 149, 149, 32, 3, 3, (0, 0)),
    ('2b 3x3p1 32->64 @147', 32, 147, 147, 64, 3, 3, (1, 1)),
    ('4a 3x3 80->192 @73', 80, 73, 73, 192, 3, 3, (0, 0)),
    ('5b 1x1s 192->208 @35', 192, 35, 35, 208, 1, 1, (0, 0)),
    ('5b 3x3 64->96 @35', 64, 35, 35, 96, 3, 3, (1, 1)),
    ('5b 5x5 48->64 @35', 48, 35, 35, 64, 5, 5, (2, 2)),
    ('6e 1x7 192->192 @17', 192, 17, 17, 192, 1, 7, (0, 3)),
    ('6c 1x1s 768->704 @17', 768, 17, 17, 704, 1, 1, (0, 0)),
    ('7c 1x1s 2048->1344 @8', 2048, 8, 8, 1344, 1, 1, (0, 0)),
    ('7c 3x3 448->384 @8', 448, 8, 8, 384, 3, 3, (1, 1)),
]


def time_layer(layer, algo, env):
    for k in ('IFCB_CONV_DEBUG', 'IFCB_CONV_MSUB', 'IFCB_CONV_ASLOTS', 'IFCB_CONV_BGROUP'):
        os.environ.pop(k, None)
    os.environ.update(env)
    name, Cin, H, W, Cout, kh, kw, pad = layer
    pb = PlanBuilder(B, dev, 'fp16')
    ipad = pad if algo == 2 else (PlanBuilder.border_for(H, W, Cout, (kh, kw), pad=pad) if algo == 0 else (0, 0))
    x = torch.randn((B, H + 2 * ipad[0], W + 2 * ipad[1], Cin), device=dev).half()
    pb.keep.append(x)
    P, Q = H + 2 * pad[0] - kh + 1, W + 2 * pad[1] - kw + 1
    out = pb.alloc(P, Q, Cout)
    w = torch.randn(Cout, Cin, kh, kw) / np.sqrt(Cin * kh * kw)
    try:
        pb.conv(View(x, pad=ipad), [dict(weight=w, scale=torch.ones(Cout), shift=torch.zeros(Cout), relu=True, out=out)],
                (1, 1), pad, algo=algo)
    except RuntimeError as e:
        return None
    for _ in range(3):
        pb.run(B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pb.run(B)
    e1.record()
    torch.cuda.synchronize()
    pb.close()
    del pb, x, out
    torch.cuda.empty_cache()
    return e0.elapsed_time(e1) / 10 * 1e3


CONFIGS = [
    ('im2col', 1, {}),
    ('im2col nomma', 1, {'IFCB_CONV_DEBUG': '2'}),
    ('im2col noepi', 1, {'IFCB_CONV_DEBUG': '4'}),
    ('im2col mma-only', 1, {'IFCB_CONV_DEBUG': '5'}),
    ('window', 2, {}),
    ('window nomma', 2, {'IFCB_CONV_DEBUG': '2'}),
    ('window noepi', 2, {'IFCB_CONV_DEBUG': '4'}),
    ('window mma-only', 2, {'IFCB_CONV_DEBUG': '5'}),
    ('window m2', 2, {'IFCB_CONV_MSUB': '2'}),
    ('pair', 3, {}),
    ('pair nomma', 3, {'IFCB_CONV_DEBUG': '2'}),
    ('pair noepi', 3, {'IFCB_CONV_DEBUG': '4'}),
    ('pair mma-only', 3, {'IFCB_CONV_DEBUG': '5'}),
    ('auto', 0, {}),
]
print('%-26s' % 'layer (us @ batch %d)' % B + ''.join('%17s' % c[0] for c in CONFIGS))
for layer in LAYERS:
    row = '%-26s' % layer[0]
    for cname, algo, env in CONFIGS:
        t = time_layer(layer, algo, env)
        row += '%17s' % ('%.1f' % t if t is not None else 'n/a')
    flops = 2.0 * B * (layer[2] + 2 * layer[7][0] - layer[5] + 1) * (layer[3] + 2 * layer[7][1] - layer[6] + 1) * layer[4] * layer[1] * layer[5] * layer[6]
    row += '   ideal@1380TF %.1f' % (flops / 1380.2e12 * 1e6)
    print(row, flush=True)
