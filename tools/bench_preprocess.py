"""K1 timing: the fused preprocess kernel on one synthetic bin, per output mode (CUDA events, L2 flushed by the output
size itself in the f32 / bf16 modes).  Prints microseconds per 512 ROIs and algorithmic GB/s (h*w bytes in + the output
tensor out) next to the measured HBM peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    from oracle import synth_bins
    from ifcb_classifier_b200 import preprocess as pp
    dev = torch.device('cuda:0')
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    R = 299
    sb = synth_bins.make_bin(0, n)
    imgs = [sb['images'][t] for t in sorted(sb['images'])]
    hs = np.array([im.shape[0] for im in imgs], np.int32)
    ws = np.array([im.shape[1] for im in imgs], np.int32)
    offs = np.concatenate([[0], np.cumsum(hs.astype(np.int64) * ws)[:-1]]).astype(np.int64)
    roi = torch.from_numpy(sb['roi']).to(dev)
    d = [torch.from_numpy(a).to(dev) for a in (offs, hs, ws)]
    peak = 6545.6
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    in_bytes = float((hs.astype(np.int64) * ws).sum())
    for mode, name, esz in ((pp.OUT_U8_GRAY, 'u8 gray plane', 1), (pp.OUT_F32_NCHW, 'f32 NCHW', 12), (pp.OUT_BF16_NCHW, 'bf16 NCHW', 6)):
        for bounds, tag in (((pp.FRAME_H, pp.FRAME_W), 'frame bound'), ((int(hs.max()), int(ws.max())), 'bin bound')):
            out = None
            for _ in range(3):
                out = pp.preprocess_rois(roi, d[0], d[1], d[2], R, out_mode=mode, out=out, max_h=bounds[0], max_w=bounds[1])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                pp.preprocess_rois(roi, d[0], d[1], d[2], R, out_mode=mode, out=out, max_h=bounds[0], max_w=bounds[1])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gb = (in_bytes + n * R * R * esz) / 1e9
            print('%-14s %-11s %d ROIs @%d: %8.1f us  = %7.1f us per 512 ROIs, %7.1f GB/s algorithmic = %.3f of %.0f GB/s'
                  % (name, tag, n, R, ms * 1e3, ms * 1e3 * 512 / n, gb / (ms * 1e-3), gb / (ms * 1e-3) / peak, peak))


if __name__ == '__main__':
    main()
