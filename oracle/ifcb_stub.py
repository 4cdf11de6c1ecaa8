"""Minimal stand-in for the parts of pyifcb the reference's hot path touches.

TEST INFRASTRUCTURE (see oracle/__init__.py).  pyifcb is absent from this image
and unpinned upstream, so this is a restatement from its published behaviour of
exactly the touch-points at ``/root/reference/neuston_data.py:14-16,446-454``,
``neuston_net.py:213-235`` and ``neuston_callbacks.py:176-181``:
``ifcb.Pid`` (``.pid .namespace .year .yearday .target with_target()``),
``ifcb.DataDirectory`` (iterates bins), a bin's ``.schema``, ``.pid``,
``.images`` (ordered mapping target -> uint8[h, w]) and ``.fileset.basepath``;
``ifcb.data.adc.SCHEMA_VERSION_1`` and ``ifcb.data.stitching.InfilledImages``.

``install()`` registers these as ``sys.modules['ifcb']`` (+ submodules) so the
reference's ``neuston_data.py`` can be imported UNMODIFIED -- used by
``tests/golden/make_golden.py`` (in the build container only).
"""
import os
import re
import sys
import types
from collections import OrderedDict
import numpy as np

SCHEMA_VERSION_1 = 'v1'
SCHEMA_VERSION_2 = 'v2'
_V2 = re.compile(r'^(D(\d{4})(\d{2})(\d{2})T(\d{6})_IFCB(\d+))(?:_(\d+))?$')
_V1 = re.compile(r'^(IFCB(\d+)_(\d{4})_(\d{3})_(\d{6}))(?:_(\d+))?$')
_COLS = {SCHEMA_VERSION_2: (15, 16, 17), SCHEMA_VERSION_1: (11, 12, 13)}  # width, height, start_byte


class Pid(object):
    def __init__(self, pid, namespace=None):
        pid = os.path.basename(str(pid))
        m2, m1 = _V2.match(pid), _V1.match(pid)
        if m2:
            self.schema_version = SCHEMA_VERSION_2
            self.bin_lid = m2.group(1)
            self.year, self.yearday = m2.group(2), m2.group(2) + m2.group(3) + m2.group(4)
            self.target = int(m2.group(7)) if m2.group(7) else None
        elif m1:
            self.schema_version = SCHEMA_VERSION_1
            self.bin_lid = m1.group(1)
            self.year, self.yearday = m1.group(3), m1.group(3) + '_' + m1.group(4)
            self.target = int(m1.group(6)) if m1.group(6) else None
        else:
            raise ValueError('invalid pid: %s' % pid)
        self.namespace = namespace

    @property
    def pid(self):
        return self.bin_lid

    def with_target(self, target):
        return '%s_%05d' % (self.bin_lid, int(target))

    def __str__(self):
        return self.bin_lid

    __repr__ = __str__


class _Fileset(object):
    def __init__(self, basepath):
        self.basepath = basepath


class FilesetBin(object):
    """One ``.adc/.hdr/.roi`` triple; ``images`` reads every ROI eagerly on first use."""

    def __init__(self, basepath):
        self.fileset = _Fileset(basepath)
        self.pid = Pid(os.path.basename(basepath))
        self.schema = self.pid.schema_version
        self._images = None

    def read_adc(self):
        path = self.fileset.basepath + '.adc'
        if os.path.getsize(path) == 0:
            return np.zeros((0, 24))
        return np.atleast_2d(np.loadtxt(path, delimiter=',', ndmin=2))

    @property
    def images(self):
        if self._images is None:
            cw, ch, cb = _COLS[self.schema]
            adc = self.read_adc()
            roi = np.fromfile(self.fileset.basepath + '.roi', dtype=np.uint8)
            out = OrderedDict()
            for i, row in enumerate(adc):
                w, h, b = int(row[cw]), int(row[ch]), int(row[cb])
                if w * h == 0:
                    continue
                out[i + 1] = roi[b:b + w * h].reshape((h, w)).copy()
            self._images = out
        return self._images

    def __len__(self):
        return len(self.images)


class DataDirectory(object):
    def __init__(self, path, whitelist=None, blacklist=None):
        self.path, self.whitelist, self.blacklist = path, whitelist, blacklist

    def __iter__(self):
        for parent, dirs, files in os.walk(self.path):
            dirs.sort()
            for f in sorted(files):
                if not f.endswith('.adc'):
                    continue
                base = os.path.join(parent, f[:-4])
                if not (os.path.isfile(base + '.roi') and os.path.isfile(base + '.hdr')):
                    continue
                name = os.path.basename(base)
                if self.whitelist is not None and not any(k in base for k in self.whitelist):
                    continue
                if self.blacklist is not None and any(k in name for k in self.blacklist):
                    continue
                try:
                    yield FilesetBin(base)
                except ValueError:
                    continue


def InfilledImages(bin):  # schema-v1 stitching is out of scope
    raise NotImplementedError('schema v1 stitched bins are not covered by the oracle')


def install():
    """Register stub ``ifcb`` modules so ``/root/reference/neuston_data.py`` imports unmodified."""
    ifcb = types.ModuleType('ifcb')
    ifcb.Pid, ifcb.DataDirectory = Pid, DataDirectory
    data = types.ModuleType('ifcb.data')
    adc = types.ModuleType('ifcb.data.adc')
    adc.SCHEMA_VERSION_1 = SCHEMA_VERSION_1
    st = types.ModuleType('ifcb.data.stitching')
    st.InfilledImages = InfilledImages
    ifcb.data, data.adc, data.stitching = data, adc, st
    sys.modules.update({'ifcb': ifcb, 'ifcb.data': data, 'ifcb.data.adc': adc, 'ifcb.data.stitching': st})
    return ifcb
