"""Synthetic IFCB bins (``.adc`` / ``.hdr`` / ``.roi`` triples) -- the workload
generator of SURVEY.md section 8(d) and the fixture source for the tests.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference ships no data; the
raw layout below restates what pyifcb (``git+https://github.com/joefutrelle/pyifcb``,
unpinned in ``/root/reference/requirements/env.dev.yml:17-18``; absent from this
image) reads at the reference call sites ``neuston_data.py:446-454``:

* ``.adc``  headerless CSV, one row per trigger, 1-based row index = ROI/target
  number; schema v2 (``DyyyymmddTHHMMSS_IFCBnnn`` bins) columns (0-based)
  ROI_X=13, ROI_Y=14, ROI_WIDTH=15, ROI_HEIGHT=16, START_BYTE=17.
* ``.roi``  headerless concatenation of 8-bit pixel blocks, ROI n =
  ``width*height`` bytes at START_BYTE, row-major ``(height, width)``.
* rows with ``width*height == 0`` carry no image and are skipped.
* ``.hdr``  free text ``key: value`` lines (unused by the hot path).

"Parity unpinned": nothing in the reference or this image pins these facts.
"""
import os
import numpy as np

ADC_COLS_V2 = 24
ROI_X, ROI_Y, ROI_WIDTH, ROI_HEIGHT, START_BYTE = 13, 14, 15, 16, 17
FRAME_W, FRAME_H = 1380, 1034


def bin_lid(bin_idx: int) -> str:
    """Deterministic schema-v2 bin id, e.g. D20260101T000000_IFCB999."""
    day = bin_idx // 1440
    minute = bin_idx % 1440
    month, dom = 1 + (day // 28) % 12, 1 + day % 28
    return 'D2026%02d%02dT%02d%02d00_IFCB999' % (month, dom, minute // 60, minute % 60)


def synth_roi_dims(rng, n):
    w = np.clip(rng.lognormal(np.log(90.0), 0.6, n), 16, FRAME_W).astype(np.int64)
    h = np.clip(rng.lognormal(np.log(60.0), 0.6, n), 16, FRAME_H).astype(np.int64)
    return h, w


def synth_roi_pixels(rng, h, w, variant='ifcb'):
    if variant == 'uniform':
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    img = rng.normal(205.0, 6.0, (h, w))
    yy, xx = np.mgrid[0:h, 0:w]
    cy, cx = rng.uniform(0.3, 0.7) * h, rng.uniform(0.3, 0.7) * w
    ry, rx = rng.uniform(0.15, 0.45) * h + 1, rng.uniform(0.15, 0.45) * w + 1
    inside = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
    img = np.where(inside, img - rng.uniform(40, 140), img)
    return np.clip(img, 0, 255).astype(np.uint8)


def make_bin(bin_idx: int, n_rois: int = 2048, variant: str = 'ifcb',
             empty_every: int = 0, dims=None):
    """Returns dict(lid, adc (float64 [rows, 24]), roi (uint8 bytes), images {target: u8[h,w]}).

    ``empty_every`` > 0 inserts a zero-sized trigger row every that many rows
    (real bins have them).  ``dims`` overrides the lognormal sizes with an
    explicit list of (h, w).
    """
    rng = np.random.default_rng(1000 + bin_idx)
    if dims is None:
        hs, ws = synth_roi_dims(rng, n_rois)
    else:
        hs = np.array([d[0] for d in dims]); ws = np.array([d[1] for d in dims])
        n_rois = len(dims)
    rows, chunks, images = [], [], {}
    pos = 0
    target = 0
    for i in range(n_rois):
        if empty_every and i % empty_every == empty_every - 1:
            target += 1
            row = np.zeros(ADC_COLS_V2); row[0] = target; row[START_BYTE] = pos
            rows.append(row)
        target += 1
        h, w = int(hs[i]), int(ws[i])
        img = synth_roi_pixels(rng, h, w, variant)
        row = np.zeros(ADC_COLS_V2)
        row[0] = target
        row[ROI_X] = int(rng.integers(0, FRAME_W - w + 1))
        row[ROI_Y] = int(rng.integers(0, FRAME_H - h + 1))
        row[ROI_WIDTH], row[ROI_HEIGHT], row[START_BYTE] = w, h, pos
        rows.append(row)
        chunks.append(img.reshape(-1))
        images[target] = img
        pos += h * w
    roi = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return dict(lid=bin_lid(bin_idx), adc=np.array(rows).reshape(-1, ADC_COLS_V2),
                roi=roi, images=images)


def write_bin(dirpath: str, b: dict) -> str:
    """Writes ``{lid}.adc/.hdr/.roi`` under ``dirpath``; returns the base path."""
    os.makedirs(dirpath, exist_ok=True)
    base = os.path.join(dirpath, b['lid'])
    with open(base + '.adc', 'w') as f:
        for row in b['adc']:
            f.write(','.join(('%d' % v) if float(v).is_integer() else repr(float(v)) for v in row) + '\n')
    with open(base + '.hdr', 'w') as f:
        f.write('softwareVersion: synthetic\nADCFileFormat: synthetic schema v2\n')
    b['roi'].tofile(base + '.roi')
    return base
