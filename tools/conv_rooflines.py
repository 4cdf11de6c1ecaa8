"""Per-layer bounds of the RUN conv kernels (Inception-v3 @299, batch 1024) beside the measured CUDA-event times: for every conv
launch of the plan, the time its tcgen05 MMA issue alone, its shared-memory port traffic alone and its HBM traffic alone would take,
and which of them binds.  CPU-only analysis of a per-layer table written by `tools/run_plan_once.py --time`.

    python tools/conv_rooflines.py profiles/r02_layer_events_final_b1024.txt [--mhz 1635] > profiles/r02_conv_rooflines.txt

Model per 128 GEMM rows (output pixels; WINDOW rows include the padded anchors the epilogue drops), constants measured on B200:
  MMA issue   taps x K16-steps x max(N/2, 32 + N/4) clk per N tile               (tools/mma_microbench.cu, mma_major_microbench.cu)
  smem port   128 B/clk for: MMA operand reads (A: 128 rows x 32 B per K16 step whatever N is; B: N x 32 B), TMA fill writes (A: once
              per tile for WINDOW [+ halo], once per tap for IM2COL; B: once per tile, shared by the m accumulators of a WINDOW tile,
              halved per SM on CTA pairs) and the epilogue's staging (write + read of the 16-bit tile)
  HBM         (input + output bytes) / (6545.6 GB/s / 148 SMs)
The algorithm and N tile of every layer come from the library (`ifcb_conv_auto_config`); m = accumulators per WINDOW tile.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifcb_classifier_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('table')
ap.add_argument('--mhz', type=float, default=1635.0, help='SM clock the table was measured at (bench line of the same box)')
ap.add_argument('--batch', type=int, default=1024)
a = ap.parse_args()

# name -> (Cin, Cout, kh, kw, stride, pad, H, W) of the fused plan (graph.py: sibling 1x1 convs of a block are one GEMM; the
# avg-pool branch's 1x1 runs before the pool)
S = {'Conv2d_2a_3x3': (32, 32, 3, 3, 1, (0, 0), 149, 149), 'Conv2d_2b_3x3': (32, 64, 3, 3, 1, (1, 1), 147, 147),
     'Conv2d_3b_1x1': (64, 80, 1, 1, 1, (0, 0), 73, 73), 'Conv2d_4a_3x3': (80, 192, 3, 3, 1, (0, 0), 73, 73)}
for blk, cin, pf in (('5b', 192, 32), ('5c', 256, 64), ('5d', 288, 64)):
    S['Mixed_%s.1x1s' % blk] = (cin, 64 + 48 + 64 + pf, 1, 1, 1, (0, 0), 35, 35)
    S['Mixed_%s.branch5x5_2' % blk] = (48, 64, 5, 5, 1, (2, 2), 35, 35)
    S['Mixed_%s.branch3x3dbl_2' % blk] = (64, 96, 3, 3, 1, (1, 1), 35, 35)
    S['Mixed_%s.branch3x3dbl_3' % blk] = (96, 96, 3, 3, 1, (1, 1), 35, 35)
S['Mixed_6a.branch3x3'] = (288, 384, 3, 3, 2, (0, 0), 35, 35)
S['Mixed_6a.branch3x3dbl_1'] = (288, 64, 1, 1, 1, (0, 0), 35, 35)
S['Mixed_6a.branch3x3dbl_2'] = (64, 96, 3, 3, 1, (1, 1), 35, 35)
S['Mixed_6a.branch3x3dbl_3'] = (96, 96, 3, 3, 2, (0, 0), 35, 35)
for blk, c7 in (('6b', 128), ('6c', 160), ('6d', 160), ('6e', 192)):
    S['Mixed_%s.1x1s' % blk] = (768, 192 + c7 + c7 + 192, 1, 1, 1, (0, 0), 17, 17)
    S['Mixed_%s.branch7x7_2' % blk] = (c7, c7, 1, 7, 1, (0, 3), 17, 17)
    S['Mixed_%s.branch7x7_3' % blk] = (c7, 192, 7, 1, 1, (3, 0), 17, 17)
    S['Mixed_%s.branch7x7dbl_2' % blk] = (c7, c7, 7, 1, 1, (3, 0), 17, 17)
    S['Mixed_%s.branch7x7dbl_3' % blk] = (c7, c7, 1, 7, 1, (0, 3), 17, 17)
    S['Mixed_%s.branch7x7dbl_4' % blk] = (c7, c7, 7, 1, 1, (3, 0), 17, 17)
    S['Mixed_%s.branch7x7dbl_5' % blk] = (c7, 192, 1, 7, 1, (0, 3), 17, 17)
S['Mixed_7a.1x1s'] = (768, 384, 1, 1, 1, (0, 0), 17, 17)
S['Mixed_7a.branch3x3_2'] = (192, 320, 3, 3, 2, (0, 0), 17, 17)
S['Mixed_7a.branch7x7x3_2'] = (192, 192, 1, 7, 1, (0, 3), 17, 17)
S['Mixed_7a.branch7x7x3_3'] = (192, 192, 7, 1, 1, (3, 0), 17, 17)
S['Mixed_7a.branch7x7x3_4'] = (192, 192, 3, 3, 2, (0, 0), 17, 17)
for blk, cin in (('7b', 1280), ('7c', 2048)):
    S['Mixed_%s.1x1s' % blk] = (cin, 320 + 384 + 448 + 192, 1, 1, 1, (0, 0), 8, 8)
    S['Mixed_%s.branch3x3_2a' % blk] = (384, 384, 1, 3, 1, (0, 1), 8, 8)
    S['Mixed_%s.branch3x3_2b' % blk] = (384, 384, 3, 1, 1, (1, 0), 8, 8)
    S['Mixed_%s.branch3x3dbl_2' % blk] = (448, 384, 3, 3, 1, (1, 1), 8, 8)
    S['Mixed_%s.branch3x3dbl_3a' % blk] = (384, 384, 1, 3, 1, (0, 1), 8, 8)
    S['Mixed_%s.branch3x3dbl_3b' % blk] = (384, 384, 3, 1, 1, (1, 0), 8, 8)

ALGO = {_lib.IFCB_CONV_IM2COL: 'im2col', _lib.IFCB_CONV_WINDOW: 'window', _lib.IFCB_CONV_IM2COL_PAIR: 'pair'}
hbm_b_per_clk = 6545.6e9 / 148 / (a.mhz * 1e6)
print(__doc__.split('Model per')[0].strip().split('\n')[0])
print('clock %.0f MHz, batch %d; times in us; bound = largest of the three\n' % (a.mhz, a.batch))
print('%-28s %-11s %4s %2s | %8s | %8s %8s %8s | %-5s %6s' % ('layer', 'algo', 'N', 'm', 'measured', 'MMA', 'smem', 'HBM', 'bound', 'meas/b'))
tot_meas = tot_bound = 0.0
for line in open(a.table):
    f = line.split()
    if len(f) < 4 or f[1] != 'conv' or f[0] not in S:
        continue
    meas = float(f[2])
    cin, cout, kh, kw, st, pad, H, W = S[f[0]]
    algo, tn = _lib.conv_auto_config(H, W, cin, cout, kh, kw, (st, st), pad)
    P, Q = (H + 2 * pad[0] - kh) // st + 1, (W + 2 * pad[1] - kw) // st + 1
    taps = kh * kw
    row_b = 64 if cin <= 32 else 128
    cblocks = 1 if cin <= 32 else (cin + 63) // 64
    last = cin - (cblocks - 1) * (row_b // 2)
    ksteps = (cblocks - 1) * (row_b // 32) + (last + 15) // 16
    n_tiles = (cout + tn - 1) // tn
    window = algo == _lib.IFCB_CONV_WINDOW
    # WINDOW layers with >= 64 output channels and 128-byte rows run on CTA pairs too (plan.cu: window_pair)
    pair = algo == _lib.IFCB_CONV_IM2COL_PAIR or (window and (cout >= 128 or (cout >= 64 and cin > 32)))
    m = 1
    if window:
        m = 4 if 4 * tn <= 256 else 2 if 2 * tn <= 256 else 1
    rows = a.batch * ((H + 2 * pad[0]) * (W + 2 * pad[1]) if window else P * Q)      # GEMM rows (WINDOW: padded anchors)
    t128 = rows / 128.0 / 148.0                                                        # 128-row tiles per SM
    mma = taps * ksteps * max(tn / 2.0, 32 + tn / 4.0) * n_tiles
    rd = taps * ksteps * (128 * 32 + tn * 32) * n_tiles
    if window:
        halo = (kh - 1) * (W + 2 * pad[1]) + kw - 1
        fill_a = cblocks * (128 * m + halo) * row_b / m * n_tiles
    else:
        fill_a = taps * cblocks * 128 * row_b * n_tiles
    fill_b = taps * cblocks * tn * row_b / m * n_tiles * (0.5 if pair else 1.0)
    epi = 2 * 128 * cout * 2
    smem = (rd + fill_a + fill_b + epi) / 128.0
    hbm = (a.batch * H * W * cin * 2 + a.batch * P * Q * cout * 2) / (rows / 128.0) / hbm_b_per_clk
    us = lambda clk: clk * t128 / a.mhz
    b = max(mma, smem, hbm)
    which = 'MMA' if b == mma else 'smem' if b == smem else 'HBM'
    tot_meas += meas
    tot_bound += us(b)
    print('%-28s %-11s %4d %2d | %8.1f | %8.1f %8.1f %8.1f | %-5s %6.2f' % (f[0], (ALGO[algo] + ('+pair' if window and pair else ''))[:11], tn, m, meas, us(mma), us(smem), us(hbm), which,
                                                                             meas / us(b)))
print('\nall %s conv launches: measured %.0f us, sum of the binding bounds %.0f us -> %.2f x' % ('listed', tot_meas, tot_bound, tot_meas / tot_bound))
