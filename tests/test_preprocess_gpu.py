"""K1 parity: CUDA preprocess vs the oracle (bit-exact, integer path) -- through the C ABI."""
import numpy as np
import pytest
import torch

from oracle.pil_resize import ref_preprocess, resize_gray_u8
from oracle import synth_bins

pytestmark = pytest.mark.gpu


def _pack(images):
    offs, hs, ws, chunks, pos = [], [], [], [], 0
    for im in images:
        offs.append(pos); hs.append(im.shape[0]); ws.append(im.shape[1])
        chunks.append(im.reshape(-1)); pos += im.size
    return np.concatenate(chunks), np.array(offs, np.int64), np.array(hs, np.int32), np.array(ws, np.int32)


def _run(cuda, images, R, img_norm, mode, lead_pad=0, pass_rule=0, **kw):
    from ifcb_classifier_b200 import preprocess as pp
    packed, offs, hs, ws = _pack(images)
    if lead_pad:   # shift every ROI so that starts are not 16-byte aligned
        packed = np.concatenate([np.zeros(lead_pad, np.uint8), packed]); offs = offs + lead_pad
    out = pp.preprocess_rois(torch.from_numpy(packed).to(cuda), torch.from_numpy(offs).to(cuda),
                             torch.from_numpy(hs).to(cuda), torch.from_numpy(ws).to(cuda), R,
                             img_norm=img_norm, out_mode=mode, pass_rule=pass_rule, **kw)
    torch.cuda.synchronize()
    return out.cpu()


EDGE_DIMS = [(8, 8), (1, 1), (1, 5), (16, 16), (60, 90), (299, 100), (100, 224), (300, 17), (224, 224),
             (299, 299), (37, 1380), (1034, 16), (1034, 1380), (500, 3), (700, 5), (1034, 2), (225, 2),
             (640, 480), (3, 1000), (223, 225), (300, 298)]


@pytest.mark.parametrize('R', [299, 224])
def test_edge_sizes_f32_bit_exact(cuda, R):
    rng = np.random.default_rng(7)
    images = [rng.integers(0, 256, d, dtype=np.uint8) for d in EDGE_DIMS]
    for norm in (None, ['0.667', '0.161'], ['0.5,0.4,0.3', '0.2,0.25,0.3']):
        got = _run(cuda, images, R, norm, 0, lead_pad=3).numpy()
        for i, im in enumerate(images):
            want = ref_preprocess(im, R, norm)
            assert np.array_equal(got[i], want), (EDGE_DIMS[i], R, norm, np.abs(got[i] - want).max())


def test_u8_gray_and_bf16_modes(cuda):
    rng = np.random.default_rng(8)
    images = [rng.integers(0, 256, d, dtype=np.uint8) for d in EDGE_DIMS]
    got = _run(cuda, images, 299, None, 2).numpy()
    for i, im in enumerate(images):
        assert np.array_equal(got[i], resize_gray_u8(im, 299)), EDGE_DIMS[i]
    got = _run(cuda, images, 224, ['0.667', '0.161'], 1).float().numpy()
    for i, im in enumerate(images):
        want = torch.from_numpy(ref_preprocess(im, 224, ['0.667', '0.161'])).bfloat16().float().numpy()
        assert np.array_equal(got[i], want), EDGE_DIMS[i]


def test_pass_rule_hv_matches_pillow8_order(cuda):
    rng = np.random.default_rng(9)
    images = [rng.integers(0, 256, d, dtype=np.uint8) for d in [(500, 3), (1034, 2), (60, 90)]]
    got = _run(cuda, images, 224, None, 2, pass_rule=1).numpy()
    for i, im in enumerate(images):
        assert np.array_equal(got[i], resize_gray_u8(im, 224, pass_rule='hv'))


def test_synthetic_bin_matches_oracle(cuda):
    b = synth_bins.make_bin(3, n_rois=256)
    images = [b['images'][t] for t in sorted(b['images'])]
    got = _run(cuda, images, 299, ['0.667', '0.161'], 0).numpy()
    for i, im in enumerate(images):
        assert np.array_equal(got[i], ref_preprocess(im, 299, ['0.667', '0.161'])), i


def test_empty_and_argument_errors(cuda):
    from ifcb_classifier_b200 import preprocess as pp
    e = torch.zeros(0, dtype=torch.int64, device=cuda)
    out = pp.preprocess_rois(torch.zeros(16, dtype=torch.uint8, device=cuda), e, e.int(), e.int(), 299)
    assert out.shape == (0, 3, 299, 299)
    one = torch.zeros(1, dtype=torch.int64, device=cuda)
    with pytest.raises(RuntimeError):
        pp.preprocess_rois(torch.zeros(16, dtype=torch.uint8, device=cuda), one, one.int() + 4, one.int() + 4,
                           299, max_h=200000, max_w=200000)


def test_tight_bounds_and_band_modes_agree(cuda):
    """The launch's shared memory follows the declared ROI bound: tight bounds (whole ROI resident), the camera-frame default
    (banded for large ROIs) and an oversized bound give identical bytes."""
    rng = np.random.default_rng(10)
    dims = [(60, 90), (16, 16), (240, 330), (130, 700), (431, 57)]
    images = [rng.integers(0, 256, d, dtype=np.uint8) for d in dims]
    want = np.stack([resize_gray_u8(im, 299) for im in images])
    for mh, mw in ((431, 700), (1034, 1380), (3000, 2000)):
        got = _run(cuda, images, 299, None, 2, max_h=mh, max_w=mw).numpy()
        assert np.array_equal(got, want), (mh, mw)


@pytest.mark.parametrize('dims,R', [((2000, 3000), 224), ((4300, 6100), 224), ((3000, 40), 299)])
def test_large_images_any_size(cuda, dims, R):
    """Photos far beyond the IFCB frame (`--type img`, TRAIN image folders): staged in bands up to the 227 KB shared-memory
    limit, beyond that the horizontal pass reads global memory directly -- same bytes as the oracle either way."""
    rng = np.random.default_rng(12)
    im = rng.integers(0, 256, dims, dtype=np.uint8)
    got = _run(cuda, [im], R, None, 2, lead_pad=5, max_h=dims[0], max_w=dims[1]).numpy()[0]
    assert np.array_equal(got, resize_gray_u8(im, R))


def test_refused_rois_are_zeroed_and_flagged(cuda):
    """A table entry that points outside the packed bytes, or a ROI larger than the declared bound, is never read or half
    processed: its output slot is zeroed (not left holding the previous batch) and the status word says why."""
    from ifcb_classifier_b200 import preprocess as pp
    rng = np.random.default_rng(13)
    good = rng.integers(0, 256, (40, 50), dtype=np.uint8)
    packed = torch.from_numpy(good.reshape(-1).copy()).to(cuda)
    offs = torch.tensor([0, 1500, 0], dtype=torch.int64, device=cuda)            # ROI 1 runs past the end of the buffer
    hs = torch.tensor([40, 40, 40], dtype=torch.int32, device=cuda)
    ws = torch.tensor([50, 50, 50], dtype=torch.int32, device=cuda)
    status = torch.zeros(1, dtype=torch.int32, device=cuda)
    out = torch.full((3, 224, 224), 7, dtype=torch.uint8, device=cuda)
    pp.preprocess_rois(packed, offs, hs, ws, 224, out_mode=2, out=out, status=status)
    want = torch.from_numpy(resize_gray_u8(good, 224))
    assert torch.equal(out[0].cpu(), want) and torch.equal(out[2].cpu(), want) and int(out[1].max()) == 0
    with pytest.raises(RuntimeError, match='outside'):
        pp.check_status(status)
    assert int(status.item()) == 0
    big = rng.integers(0, 256, (900, 1200), dtype=np.uint8)                       # declared bound is far smaller
    out = torch.full((1, 3, 224, 224), 7.0, device=cuda)
    pp.preprocess_rois(torch.from_numpy(big.reshape(-1).copy()).to(cuda), offs[:1], hs[:1] * 0 + 900, ws[:1] * 0 + 1200, 224, out=out,
                       max_h=64, max_w=64, status=status)
    assert float(out.abs().max()) == 0.0
    with pytest.raises(RuntimeError, match='larger'):
        pp.check_status(status)
