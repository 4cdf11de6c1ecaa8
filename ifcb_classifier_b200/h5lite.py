"""Minimal HDF5 writer / reader for the reference's fixed result-file schemas (no h5py / libhdf5 needed).

The reference's DEFAULT RUN output is ``D{BIN_YEAR}/D{BIN_DATE}/{BIN_ID}_class.h5`` (neuston_net.py:180-182) written by
``_save_run_results_hdf`` (neuston_callbacks.py:252-268), and ``SaveValidationResults`` can write ``.h5`` too
(neuston_callbacks.py:139-156).  h5py is not installed in this image, so the subset of the HDF5 file format those two
writers produce is emitted directly (HDF5 File Format Specification, version 0 superblock -- what libhdf5 writes by
default with ``libver='earliest'``):

  * one root group: version-1 object header with a symbol-table message, one v1 B-tree node ("TREE", type 0) with one
    symbol-table node ("SNOD"), names in a local heap ("HEAP");
  * datasets: version-1 object headers (dataspace, datatype, fill value, layout [, filter pipeline] [, attributes]);
    numeric and variable-length-string datasets are CHUNKED with ONE chunk = the whole array and the deflate filter
    (``compression='gzip'`` in h5py), indexed by a v1 chunk B-tree ("TREE", type 1);
  * ``h5py.Empty('f')`` = null dataspace (dataspace message version 2, type 2), float32, contiguous layout without storage;
  * variable-length UTF-8 strings (``h5py.string_dtype()``): 16-byte descriptors (length, global-heap address, index)
    pointing into global heap collections ("GCOL"); python ``str`` attributes are scalar variable-length strings too;
  * float16 = HDF5 floating-point class with a 5-bit exponent at bit 10, bias 15 (what h5py maps ``'float16'`` to).

``read`` parses the same subset back (used by the round-trip tests and by anyone who needs to read the files where no
HDF5 library is installed).  NOTE: nothing in this image can open these files with libhdf5, so cross-validation against
the C library is not possible here; the writer follows the specification field by field and the reader is an independent
re-parse (every address, size and B-tree key is checked on the way).
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b'\x89HDF\r\n\x1a\n'
_LEAF_K, _INTERNAL_K, _CHUNK_K = 32, 16, 32          # group leaf K (<= 64 entries in the one SNOD), group internal K, chunk B-tree K


class Empty(object):
    """``h5py.Empty(dtype)``: a dataset with a null dataspace."""

    def __init__(self, dtype='f'):
        self.dtype = np.dtype(dtype)


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


# ---- datatype messages ---------------------------------------------------------------------------------------
def _dt_float(size):
    exp_loc, exp_bits, man_bits, bias = {2: (10, 5, 10, 15), 4: (23, 8, 23, 127), 8: (52, 11, 52, 1023)}[size]
    return struct.pack('<BBBBI', 0x11, 0x20, 8 * size - 1, 0, size) + \
        struct.pack('<HHBBBBI', 0, 8 * size, exp_loc, exp_bits, 0, man_bits, bias)


def _dt_int(size, signed):
    return struct.pack('<BBBBI', 0x10, 0x08 if signed else 0x00, 0, 0, size) + struct.pack('<HH', 0, 8 * size)


def _dt_vlen_str():
    base = struct.pack('<BBBBI', 0x13, 0x00, 0, 0, 1)                       # H5T_C_S1: 1-byte null-terminated ASCII string
    return struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) + base            # variable-length, type = string, charset UTF-8


def _dt_for(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == 'f':
        return _dt_float(dtype.itemsize)
    if dtype.kind in 'iu':
        return _dt_int(dtype.itemsize, dtype.kind == 'i')
    raise TypeError('h5lite: unsupported dtype %s' % dtype)


def _dataspace(shape):
    if shape is None:                                                       # null dataspace (version 2, type 2)
        return struct.pack('<BBBB', 2, 0, 0, 2)
    return struct.pack('<BBBB4x', 1, len(shape), 0, 0) + b''.join(struct.pack('<Q', int(d)) for d in shape)


class _Writer(object):
    def __init__(self):
        self.buf = bytearray()

    def tell(self):
        return len(self.buf)

    def align(self, n=8):
        self.buf += b'\0' * (-len(self.buf) % n)

    def put(self, data):
        self.align()
        addr = len(self.buf)
        self.buf += data
        return addr

    # -- global heap: variable-length strings -> list of (length, collection address, index) --
    def preload_strings(self, strings):
        """Places ALL variable-length strings of a file into shared global-heap collections up front (one 4 KB collection
        per attribute would waste 16 KB per result file); ``put_strings`` then hands the descriptors out in order."""
        enc = [s if isinstance(s, bytes) else str(s).encode('utf-8') for s in strings]
        self._pool = {}
        for s, d in zip(enc, self.put_strings(enc)):
            self._pool.setdefault(s, []).append(d)

    def put_strings(self, strings):
        out = []
        enc = [s if isinstance(s, bytes) else str(s).encode('utf-8') for s in strings]
        pool = getattr(self, '_pool', None)
        if pool is not None and all(pool.get(s) for s in set(enc)) and all(len(pool[s]) >= enc.count(s) for s in set(enc)):
            return [pool[s].pop(0) for s in enc]
        for c0 in range(0, len(enc), 60000):                               # object indices are 16 bit
            part = enc[c0:c0 + 60000]
            body, idx = bytearray(), []
            for s in part:
                if len(s) == 0:                                            # empty string: null descriptor, no heap object
                    idx.append(0)
                    continue
                idx.append(1 + sum(1 for j in idx if j))
                body += struct.pack('<HHIQ', idx[-1], 1, 0, len(s)) + _pad8(s)
            size = 16 + len(body)
            total = max(4096, (size + 16 + 4095) // 4096 * 4096)           # room for the free-space object (index 0)
            free = total - size
            body += struct.pack('<HHIQ', 0, 0, 0, free) + b'\0' * (free - 16)
            addr = self.put(b'GCOL' + struct.pack('<B3xQ', 1, total) + bytes(body))
            out.extend((len(s), addr if i else 0, i) for s, i in zip(part, idx))
        return out

    def vlen_descriptors(self, strings):
        return b''.join(struct.pack('<IQI', n, a, i) for n, a, i in self.put_strings(strings))

    # -- object header, version 1 --
    def object_header(self, messages):
        body = bytearray()
        for mtype, data, flags in messages:
            data = _pad8(data)
            body += struct.pack('<HHB3x', mtype, len(data), flags) + data
        return self.put(struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body)) + bytes(body))

    def attribute(self, name, value):
        nm = name.encode('utf-8') + b'\0'
        if isinstance(value, (str, bytes)):
            dt, data = _dt_vlen_str(), self.vlen_descriptors([value])
        elif isinstance(value, (bool, np.bool_)):
            dt, data = _dt_int(1, True), struct.pack('<b', int(value))
        elif isinstance(value, (int, np.integer)):
            dt, data = _dt_int(8, True), struct.pack('<q', int(value))
        elif isinstance(value, (float, np.floating)):
            dt, data = _dt_float(8), struct.pack('<d', float(value))
        else:
            raise TypeError('h5lite: attribute %s has unsupported type %s' % (name, type(value).__name__))
        ds = _dataspace(())
        return (0x000C, struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data, 0)

    def dataset(self, value, dtype=None, attrs=None, compression='gzip', level=4):
        msgs = []
        if isinstance(value, Empty):
            msgs.append((0x0001, _dataspace(None), 0))
            msgs.append((0x0003, _dt_for(value.dtype), 1))
            msgs.append((0x0005, struct.pack('<BBBBI', 2, 2, 2, 1, 0), 0))
            msgs.append((0x0008, struct.pack('<BBQQ', 3, 1, UNDEF, 0), 0))
        else:
            if dtype == 'vlen_str':
                arr = np.asarray(value, dtype=object)
                shape = arr.shape
                raw, dt, esize = self.vlen_descriptors(list(arr.reshape(-1))), _dt_vlen_str(), 16
            else:
                arr = np.ascontiguousarray(np.asarray(value).astype(np.dtype(dtype) if dtype is not None else np.asarray(value).dtype))
                arr = arr.astype(arr.dtype.newbyteorder('<'), copy=False)
                shape = arr.shape
                raw, dt, esize = arr.tobytes(), _dt_for(arr.dtype), arr.dtype.itemsize
            if len(shape) == 0 or any(int(d) == 0 for d in shape):
                raise ValueError('h5lite: chunked datasets need rank >= 1 and non-empty dimensions (shape %s)' % (shape,))
            rank = len(shape)
            gz = compression == 'gzip'
            stored = zlib.compress(raw, level) if gz else raw
            chunk_addr = self.put(stored)
            # v1 B-tree leaf with one chunk; the node is allocated at its full (2K children, 2K+1 keys) size
            key = lambda nbytes, offs: struct.pack('<II', nbytes, 0) + b''.join(struct.pack('<Q', int(o)) for o in offs)
            node = b'TREE' + struct.pack('<BBHQQ', 1, 0, 1, UNDEF, UNDEF)
            node += key(len(stored), [0] * (rank + 1)) + struct.pack('<Q', chunk_addr) + key(0, list(shape) + [esize])
            ksize = 8 + 8 * (rank + 1)
            node += b'\0' * (24 + 2 * _CHUNK_K * 8 + (2 * _CHUNK_K + 1) * ksize - len(node))
            btree = self.put(node)
            msgs.append((0x0001, _dataspace(shape), 0))
            msgs.append((0x0003, dt, 1))
            msgs.append((0x0005, struct.pack('<BBBBI', 2, 3, 2, 1, 0), 0))
            if gz:
                msgs.append((0x000B, struct.pack('<BB2x4x', 1, 1) + struct.pack('<HHHH', 1, 8, 1, 1) + b'deflate\0' +
                             struct.pack('<I4x', level), 0))
            msgs.append((0x0008, struct.pack('<BBBQ', 3, 2, rank + 1, btree) +
                         b''.join(struct.pack('<I', int(d)) for d in list(shape) + [esize]), 0))
        for k, v in (attrs or {}).items():
            msgs.append(self.attribute(k, v))
        return self.object_header(msgs)


def write(path, datasets):
    """``datasets``: ordered mapping name -> dict(data=..., dtype=None | numpy dtype | 'vlen_str', attrs={...},
    compression='gzip' | None).  ``data`` may be ``Empty(dtype)``."""
    w = _Writer()
    w.buf += b'\0' * 96                                                     # superblock, patched at the end
    names = sorted(datasets, key=lambda s: s.encode('utf-8'))              # symbol-table nodes are sorted by name
    if len(names) > 2 * _LEAF_K:
        raise ValueError('h5lite: at most %d datasets per file' % (2 * _LEAF_K))
    strings = []
    for n in names:
        d = datasets[n]
        if d.get('dtype') == 'vlen_str':
            strings.extend(np.asarray(d['data'], dtype=object).reshape(-1))
        strings.extend(v for v in (d.get('attrs') or {}).values() if isinstance(v, (str, bytes)))
    if strings:
        w.preload_strings(strings)
    addrs = {}
    for n in names:
        d = datasets[n]
        addrs[n] = w.dataset(d['data'], d.get('dtype'), d.get('attrs'), d.get('compression', 'gzip'))
    # local heap: "" at offset 0, then the names, then one free block
    heap, offs = bytearray(8), {}
    for n in names:
        offs[n] = len(heap)
        heap += _pad8(n.encode('utf-8') + b'\0')
    free_off = len(heap)
    heap += struct.pack('<QQ', 1, 32) + b'\0' * 16                          # free block: next = H5HL_FREE_NULL (1), size 32
    heap_data = w.put(bytes(heap))
    heap_addr = w.put(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free_off, heap_data))
    snod = b'SNOD' + struct.pack('<BBH', 1, 0, len(names))
    for n in names:
        snod += struct.pack('<QQII16x', offs[n], addrs[n], 0, 0)
    snod += b'\0' * (8 + 2 * _LEAF_K * 40 - len(snod))
    snod_addr = w.put(snod)
    node = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1 if names else 0, UNDEF, UNDEF)
    node += struct.pack('<QQQ', 0, snod_addr, offs[names[-1]] if names else 0)
    node += b'\0' * (24 + 2 * _INTERNAL_K * 8 + (2 * _INTERNAL_K + 1) * 8 - len(node))
    btree_addr = w.put(node)
    root = w.object_header([(0x0011, struct.pack('<QQ', btree_addr, heap_addr), 0)])
    w.align()
    eof = w.tell()
    sb = _SIG + struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack('<HHI', _LEAF_K, _INTERNAL_K, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
    sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', btree_addr, heap_addr)
    assert len(sb) == 96
    w.buf[0:96] = sb
    with open(path, 'wb') as f:
        f.write(w.buf)
    return path


# =====================================================================================================================
# reader (the same subset)
# =====================================================================================================================
class Dataset(object):
    def __init__(self, data, attrs, shape, dtype, filters):
        self.data, self.attrs, self.shape, self.dtype, self.filters = data, attrs, shape, dtype, filters

    def __getitem__(self, key):
        return self.data[key]


class _Reader(object):
    def __init__(self, buf):
        self.b = buf

    def u(self, fmt, off):
        return struct.unpack_from('<' + fmt, self.b, off)

    def datatype(self, off):
        cv, b0, b1, b2, size = self.u('BBBBI', off)
        cls, ver = cv & 15, cv >> 4
        assert ver == 1, 'datatype version %d' % ver
        if cls == 0:
            assert b0 & 1 == 0, 'big-endian integers'
            _, prec = self.u('HH', off + 8)
            assert prec == 8 * size
            return np.dtype('%s%d' % ('i' if b0 & 8 else 'u', size)), 12
        if cls == 1:
            _, prec, eloc, ebits, mloc, mbits, bias = self.u('HHBBBBI', off + 8)
            want = {2: (16, 10, 5, 0, 10, 15), 4: (32, 23, 8, 0, 23, 127), 8: (64, 52, 11, 0, 52, 1023)}[size]
            assert (prec, eloc, ebits, mloc, mbits, bias) == want and b1 == 8 * size - 1 and b0 == 0x20, 'non-IEEE float layout'
            return np.dtype('f%d' % size), 20
        if cls == 9:
            assert b0 & 15 == 1 and size == 16, 'only variable-length strings'
            base, n = self.datatype(off + 8)
            assert base == 'str1'
            return 'vlen_str', 8 + n
        if cls == 3:
            assert size == 1
            return 'str1', 8
        raise TypeError('datatype class %d' % cls)

    def dataspace(self, off):
        ver = self.b[off]
        if ver == 2:
            _, rank, flags, typ = self.u('BBBB', off)
            if typ == 2:
                return None
            return tuple(self.u('Q' * rank, off + 4))
        assert ver == 1
        _, rank, flags, _ = self.u('BBBB', off)
        return tuple(self.u('Q' * rank, off + 8))

    def gheap(self, addr, index):
        assert self.b[addr:addr + 4] == b'GCOL', 'bad global heap signature'
        ver, total = self.u('B3xQ', addr + 4)
        p, end = addr + 16, addr + total
        while p + 16 <= end:
            idx, ref, _, size = self.u('HHIQ', p)
            if idx == index:
                return bytes(self.b[p + 16:p + 16 + size])
            if idx == 0:
                break
            p += 16 + (size + 7) // 8 * 8
        raise KeyError('global heap object %d not found' % index)

    def vlen(self, raw, count):
        out = []
        for i in range(count):
            n, addr, idx = struct.unpack_from('<IQI', raw, 16 * i)
            s = self.gheap(addr, idx) if n else b''
            assert len(s) == n
            out.append(s.decode('utf-8'))
        return out

    def messages(self, addr):
        ver, _, nmsg, refc, hsize = self.u('BBHII', addr)
        assert ver == 1, 'object header version %d' % ver
        p, end, out = addr + 16, addr + 16 + hsize, []
        while p < end and len(out) < nmsg:
            mtype, msize, flags = self.u('HHB', p)
            out.append((mtype, p + 8, msize))
            p += 8 + msize
        assert len(out) == nmsg and p <= end
        return out

    def object(self, addr):
        shape = dtype = layout = None
        filters, attrs = [], {}
        for mtype, p, n in self.messages(addr):
            if mtype == 0x0001:
                shape = self.dataspace(p)
            elif mtype == 0x0003:
                dtype, _ = self.datatype(p)
            elif mtype == 0x0008:
                ver, cls = self.u('BB', p)
                assert ver == 3
                if cls == 1:
                    layout = ('contiguous',) + self.u('QQ', p + 2)
                elif cls == 2:
                    nd, bt = self.u('BQ', p + 2)
                    layout = ('chunked', bt, self.u('I' * nd, p + 11))
                else:
                    raise TypeError('layout class %d' % cls)
            elif mtype == 0x000B:
                ver, nf = self.u('BB', p)
                assert ver == 1
                q = p + 8
                for _ in range(nf):
                    fid, nlen, fl, ncd = self.u('HHHH', q)
                    q += 8 + (nlen + 7) // 8 * 8
                    cd = self.u('I' * ncd, q)
                    q += 4 * ncd + (4 if ncd % 2 else 0)
                    filters.append((fid, cd))
            elif mtype == 0x000C:
                ver, _, nlen, dlen, slen = self.u('BBHHH', p)
                assert ver == 1
                q = p + 8
                name = bytes(self.b[q:q + nlen]).rstrip(b'\0').decode('utf-8')
                q += (nlen + 7) // 8 * 8
                adt, _ = self.datatype(q)
                q += (dlen + 7) // 8 * 8
                ashape = self.dataspace(q)
                q += (slen + 7) // 8 * 8
                assert ashape == (), 'only scalar attributes'
                if adt == 'vlen_str':
                    attrs[name] = self.vlen(self.b[q:q + 16], 1)[0]
                else:
                    attrs[name] = np.frombuffer(self.b, adt, 1, q)[0].item()
        if shape is None:
            return Dataset(Empty(dtype), attrs, None, dtype, filters)
        assert layout[0] == 'chunked', layout
        _, bt, cdims = layout
        rank = len(shape)
        assert tuple(cdims[:-1]) == tuple(shape), 'one chunk = the whole array'
        assert self.b[bt:bt + 4] == b'TREE'
        ntype, level, used, left, right = self.u('BBHQQ', bt + 4)
        assert (ntype, level, used, left, right) == (1, 0, 1, UNDEF, UNDEF)
        nbytes, mask = self.u('II', bt + 24)
        offs = self.u('Q' * (rank + 1), bt + 32)
        assert all(o == 0 for o in offs) and mask == 0
        child, = self.u('Q', bt + 32 + 8 * (rank + 1))
        raw = bytes(self.b[child:child + nbytes])
        for fid, cd in filters:
            assert fid == 1, 'filter %d' % fid
            raw = zlib.decompress(raw)
        count = int(np.prod(shape))
        if dtype == 'vlen_str':
            assert cdims[-1] == 16 and len(raw) == 16 * count
            data = np.asarray(self.vlen(raw, count), dtype=object).reshape(shape)
        else:
            assert cdims[-1] == dtype.itemsize and len(raw) == dtype.itemsize * count
            data = np.frombuffer(raw, dtype).reshape(shape).copy()
        return Dataset(data, attrs, tuple(shape), dtype, filters)


def read(path):
    """-> {name: Dataset(data, attrs, shape, dtype, filters)} for a file written by ``write`` (or by libhdf5 using the same
    subset of the format)."""
    with open(path, 'rb') as f:
        buf = f.read()
    r = _Reader(memoryview(buf))
    assert bytes(buf[:8]) == _SIG, 'not an HDF5 file'
    ver, fs, rg, _, sh, so, sl, _ = r.u('BBBBBBBB', 8)
    assert (ver, so, sl) == (0, 8, 8), 'superblock version %d' % ver
    leaf_k, int_k, flags = r.u('HHI', 16)
    base, _, eof, _ = r.u('QQQQ', 24)
    assert base == 0 and eof == len(buf), 'end-of-file address %d != file size %d' % (eof, len(buf))
    name_off, root, cache, _ = r.u('QQII', 56)
    bt = heap = None
    for mtype, p, n in r.messages(root):
        if mtype == 0x0011:
            bt, heap = r.u('QQ', p)
    assert bt is not None, 'root group has no symbol table'
    if cache == 1:
        assert r.u('QQ', 80) == (bt, heap)
    assert bytes(buf[heap:heap + 4]) == b'HEAP'
    _, hsize, hfree, hdata = r.u('B3xQQQ', heap + 4)
    assert bytes(buf[bt:bt + 4]) == b'TREE'
    ntype, level, used, left, right = r.u('BBHQQ', bt + 4)
    assert (ntype, level) == (0, 0)
    out = {}
    for e in range(used):
        snod, = r.u('Q', bt + 24 + 8 + 16 * e)
        assert bytes(buf[snod:snod + 4]) == b'SNOD'
        _, _, nsym = r.u('BBH', snod + 4)
        prev = b''
        for i in range(nsym):
            noff, oaddr, ctype, _ = r.u('QQII', snod + 8 + 40 * i)
            assert noff < hsize
            end = buf.index(b'\0', hdata + noff)
            name = bytes(buf[hdata + noff:end])
            assert name > prev, 'symbol table entries must be sorted'
            prev = name
            out[name.decode('utf-8')] = r.object(oaddr)
    return out
