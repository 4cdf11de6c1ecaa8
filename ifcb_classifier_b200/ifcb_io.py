"""Host-side IFCB bin ingest for the RUN path (replaces the pyifcb touch-points of
reference neuston_data.py:446-454 and neuston_net.py:213-235).

A bin is an ``.adc`` / ``.hdr`` / ``.roi`` triple.  Instead of materialising one
ndarray per ROI (what ``bin.images.items()`` does), ``RawBin`` keeps the ``.roi``
file as ONE byte buffer plus an (offset, height, width) table parsed from the
``.adc`` -- exactly the layout the preprocess kernel consumes, so no per-ROI host
work remains.  pyifcb is unpinned upstream and absent here; the schema facts
(column indices, 1-based target numbers, zero-area rows skipped, ``(h, w)``
row-major blocks) are restated in DESIGN.md and marked "parity unpinned".
"""
import os
import re

import numpy as np

SCHEMA_VERSION_1 = 'v1'
SCHEMA_VERSION_2 = 'v2'
# (ROI_WIDTH, ROI_HEIGHT, START_BYTE) column indices per ADC schema
_ADC_COLS = {SCHEMA_VERSION_2: (15, 16, 17), SCHEMA_VERSION_1: (11, 12, 13)}
_RE_V2 = re.compile(r'^(D(\d{4})(\d{2})(\d{2})T(\d{6})_IFCB(\d+))(?:_(\d+))?$')
_RE_V1 = re.compile(r'^(IFCB(\d+)_(\d{4})_(\d{3})_(\d{6}))(?:_(\d+))?$')


class Pid(object):
    """Bin / target identifier with the attributes the reference uses:
    ``.pid .bin_lid .namespace .year .yearday .target`` and ``with_target()``."""

    def __init__(self, pid, namespace=None):
        name = os.path.basename(str(pid))
        m = _RE_V2.match(name)
        if m:
            self.schema_version = SCHEMA_VERSION_2
            self.year, self.yearday = m.group(2), m.group(2) + m.group(3) + m.group(4)
            tgt = m.group(7)
        else:
            m = _RE_V1.match(name)
            if not m:
                raise ValueError('invalid IFCB pid: %s' % pid)
            self.schema_version = SCHEMA_VERSION_1
            self.year, self.yearday = m.group(3), m.group(3) + '_' + m.group(4)
            tgt = m.group(6)
        self.bin_lid = m.group(1)
        self.target = int(tgt) if tgt else None
        self.namespace = namespace

    @property
    def pid(self):
        return self.bin_lid

    def with_target(self, target):
        return '%s_%05d' % (self.bin_lid, int(target))

    def __str__(self):
        return self.bin_lid

    __repr__ = __str__


def parse_adc(path, schema=SCHEMA_VERSION_2):
    """``.adc`` CSV -> (targets int32[n], offsets int64[n], heights int32[n], widths int32[n]);
    rows with zero area are dropped, target = 1-based row number.  Parsed by the library's C++
    one-pass parser (``ifcb_parse_adc``; the call releases the GIL, so ingest threads overlap)."""
    from . import _lib
    cw, ch, cb = _ADC_COLS[schema]
    with open(path, 'rb') as f:
        buf = f.read()
    cap = buf.count(b'\n') + 1
    targets = np.empty(cap, np.int32)
    offsets = np.empty(cap, np.int64)
    heights = np.empty(cap, np.int32)
    widths = np.empty(cap, np.int32)
    n = _lib.lib().ifcb_parse_adc(buf, len(buf), cw, ch, cb, cap, targets.ctypes.data, offsets.ctypes.data, heights.ctypes.data,
                                  widths.ctypes.data)
    if n < 0:
        _lib.check(-1, 'parse_adc(%s)' % path)
    return targets[:n].copy(), offsets[:n].copy(), heights[:n].copy(), widths[:n].copy()


class RawBin(object):
    """One bin as the GPU path wants it: raw ``.roi`` bytes + ROI table."""

    def __init__(self, basepath=None, pid=None, roi=None, targets=None, offsets=None, heights=None, widths=None,
                 into=None, orientation='hw'):
        """``into``: optional ``f(nbytes) -> writable uint8 ndarray`` (>= nbytes): the ``.roi`` file is read straight into
        it (``readinto``) -- the RUN driver passes slots of a pinned ring so that the upload is an asynchronous DMA.
        ``orientation``: how a ROI's ``ROI_WIDTH x ROI_HEIGHT`` byte block is laid out (pyifcb is absent and unpinned
        upstream): 'hw' = ROI_HEIGHT rows of ROI_WIDTH bytes (default; what oracle/ifcb_stub.py restates),
        'wh' = ROI_WIDTH rows of ROI_HEIGHT bytes (the table's two columns swap roles; no pixel is moved)."""
        if orientation not in ('hw', 'wh'):
            raise ValueError("orientation must be 'hw' or 'wh'")
        self.orientation = orientation
        if basepath is not None:
            self.basepath = basepath
            self.pid = Pid(os.path.basename(basepath))
            self.schema = self.pid.schema_version
            if self.schema == SCHEMA_VERSION_1:
                raise NotImplementedError('schema v1 (stitched) bins are not supported yet')
            self.targets, self.offsets, self.heights, self.widths = parse_adc(basepath + '.adc', self.schema)
            if into is None:
                self.roi = np.fromfile(basepath + '.roi', dtype=np.uint8)
            else:
                size = os.path.getsize(basepath + '.roi')
                buf = into(size)
                with open(basepath + '.roi', 'rb', buffering=0) as f:
                    got = f.readinto(memoryview(buf)[:size])
                    while got < size:
                        more = f.readinto(memoryview(buf)[got:size])
                        if not more:
                            break
                        got += more
                self.roi = buf[:got]
        else:
            self.basepath = None
            self.pid = pid if isinstance(pid, Pid) else Pid(pid)
            self.schema = self.pid.schema_version
            self.roi, self.targets, self.offsets = roi, targets, offsets
            self.heights, self.widths = heights, widths
        if orientation == 'wh':
            self.heights, self.widths = self.widths, self.heights
        end = self.offsets + self.heights.astype(np.int64) * self.widths.astype(np.int64)
        if len(end) and (end.max() > self.roi.size or self.offsets.min() < 0):
            raise ValueError('%s: ADC table points outside the .roi file' % self.pid)

    def __len__(self):
        return int(self.targets.shape[0])

    @property
    def pids(self):
        return [self.pid.with_target(t) for t in self.targets]

    def image(self, i):
        """uint8[h, w] view of ROI i (host side; for inspection / tests)."""
        o, h, w = int(self.offsets[i]), int(self.heights[i]), int(self.widths[i])
        return self.roi[o:o + h * w].reshape(h, w)


class DataDirectory(object):
    """Walks a directory tree yielding ``RawBin`` for every complete triple, in sorted
    order (deterministic -- the multi-GPU sharding relies on it)."""

    def __init__(self, path, whitelist=None, blacklist=None):
        self.path, self.whitelist, self.blacklist = path, whitelist, blacklist

    def basepaths(self):
        out = []
        for parent, dirs, files in os.walk(self.path):
            dirs.sort()
            for f in sorted(files):
                if not f.endswith('.adc'):
                    continue
                base = os.path.join(parent, f[:-4])
                if not (os.path.isfile(base + '.roi') and os.path.isfile(base + '.hdr')):
                    continue
                name = os.path.basename(base)
                if not (_RE_V2.match(name) or _RE_V1.match(name)):
                    continue
                if self.whitelist is not None and not any(k in base for k in self.whitelist):
                    continue
                if self.blacklist is not None and any(k in name for k in self.blacklist):
                    continue
                out.append(base)
        return out

    def __iter__(self):
        for base in self.basepaths():
            yield RawBin(base)
