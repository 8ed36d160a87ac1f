"""Dependency-free reader (and fixture writer) for the slice of HDF5 that BehaveNet data files use.

The reference keeps every session in one ``data.hdf5``: a group per signal (``images``, ``masks``, ``labels``,
``neural`` ...) holding one dataset per trial, ``trial_%04i`` (reference ``behavenet/data/data_generator.py:253-303``,
written as in ``docs/source/data_structure.rst:59-79`` / ``behavenet/data/preprocess.py:80``: plain
``create_dataset(name, data=array, dtype=...)`` calls, i.e. contiguous, unfiltered datasets of fixed-point or
floating-point numbers).  The reference reads them through ``h5py``; this module reads the same bytes with
``struct`` + ``numpy`` so that the input pipeline (``HDF5Source``) does not depend on an HDF5 library being
installed, and exposes the handful of ``h5py`` calls the reference makes (``File(path, 'r', ...)`` as a context
manager, ``f[group][name]``, ``len(group)``, ``keys()``, ``dataset.shape`` / ``.dtype``, ``dataset[()]`` and
leading-axis slices).

Supported on-disk structures (HDF5 file-format specification, version 3.0):

* superblock versions 0/1 (the library default) and 2/3 (``libver='latest'``), with a user block / base address;
* groups stored as symbol tables (version-1 B-tree + local heap + symbol-table nodes) and "new style" groups with
  compact link messages or dense link storage (fractal heap; the name-index B-tree is not needed to enumerate);
* version-1 and version-2 object headers with continuation blocks;
* dataspace versions 1/2, fixed-point and floating-point datatypes of either byte order;
* data layout versions 1-4 classes compact and contiguous, and version 1-3 chunked layout (version-1 B-tree chunk
  index) with the deflate, shuffle and fletcher32 filters.  Version-4 chunk indexes (single-chunk / implicit /
  fixed-array / extensible-array / v2 B-tree) raise ``NotImplementedError``.

Pinning.  The old-style path (superblock 0, symbol-table groups, version-1 object headers, contiguous layout) is
checked against a file written by the HDF5 library itself (a MATLAB v7.3 file that ships with scipy's test data;
tests/test_hdf5_lite.py).  The ``libver='latest'`` structures are implemented from the published specification only:
no file of that kind exists in this image and none can be produced without the library, so that branch is
UNPINNED and says so when it is taken (``File.unpinned_format``).  ``write`` emits old-style files (the pinned
format) and exists for fixtures and tests.
"""

import mmap
import struct
import zlib

import numpy as np

__all__ = ['File', 'Group', 'Dataset', 'write']

_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class _Reader:
    """Little-endian field reader over the mapped file; addresses are relative to the base address."""

    def __init__(self, buf, base):
        self.buf, self.base, self.osz, self.lsz = buf, base, 8, 8

    def u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], 'little')

    def addr(self, off):
        v = self.u(off, self.osz)
        return None if v == (1 << (8 * self.osz)) - 1 else v + self.base

    def length(self, off):
        return self.u(off, self.lsz)

    def addr_bytes(self, b, o):
        """Address field inside an already extracted message body."""
        v = int.from_bytes(b[o:o + self.osz], 'little')
        return None if v == (1 << (8 * self.osz)) - 1 else v + self.base


def _parse_datatype(b):
    cls, ver = b[0] & 0x0F, b[0] >> 4
    bits0 = b[1]
    size = int.from_bytes(b[4:8], 'little')
    order = '>' if bits0 & 1 else '<'
    if cls == 0:
        kind = 'i' if bits0 & 0x08 else 'u'
    elif cls == 1:
        kind = 'f'
    else:
        raise NotImplementedError('HDF5 datatype class %d (version %d) is not a plain number' % (cls, ver))
    if size not in (1, 2, 4, 8) or (kind == 'f' and size == 1):
        raise NotImplementedError('HDF5 %s datatype of %d bytes' % ('float' if kind == 'f' else 'integer', size))
    return np.dtype('%s%s%d' % ('|' if size == 1 else order, kind, size))


def _parse_dataspace(r, b):
    ver, rank, flags = b[0], b[1], b[2]
    if ver == 1:
        o = 8
    elif ver == 2:
        if b[3] == 2:
            raise NotImplementedError('null dataspace')
        o = 4
    else:
        raise NotImplementedError('dataspace message version %d' % ver)
    return tuple(int.from_bytes(b[o + i * r.lsz:o + (i + 1) * r.lsz], 'little') for i in range(rank))


def _parse_filters(b):
    ver, n = b[0], b[1]
    o = 8 if ver == 1 else 2
    out = []
    for _ in range(n):
        fid = int.from_bytes(b[o:o + 2], 'little')
        o += 2
        if ver == 1 or fid >= 256:
            nlen = int.from_bytes(b[o:o + 2], 'little')
            o += 2
        else:
            nlen = 0
        ncd = int.from_bytes(b[o + 2:o + 4], 'little')
        o += 4
        o += (nlen + 7) // 8 * 8 if ver == 1 else nlen
        cd = [int.from_bytes(b[o + 4 * i:o + 4 * i + 4], 'little') for i in range(ncd)]
        o += 4 * ncd
        if ver == 1 and ncd % 2:
            o += 4
        out.append((fid, cd))
    return out


class _Object:
    """One object header, decoded into its messages: [(type, bytes)]."""

    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        r, buf = f._r, f._r.buf
        self.msgs = []
        if buf[addr:addr + 4] == b'OHDR':
            f.unpinned_format = True
            self._v2(r, buf, addr)
        else:
            self._v1(r, buf, addr)

    def _v1(self, r, buf, addr):
        if buf[addr] != 1:
            raise ValueError('object header version %d at %d' % (buf[addr], addr))
        nmsg = r.u(addr + 2, 2)
        blocks = [(addr + 16, r.u(addr + 8, 4))]
        while blocks and len(self.msgs) < nmsg:
            o, size = blocks.pop(0)
            end = o + size
            while o + 8 <= end and len(self.msgs) < nmsg:
                mtype, msize = r.u(o, 2), r.u(o + 2, 2)
                body = bytes(buf[o + 8:o + 8 + msize])
                o += 8 + msize
                if mtype == 0x10:
                    blocks.append((r.addr_bytes(body, 0), int.from_bytes(body[r.osz:r.osz + r.lsz], 'little')))
                self.msgs.append((mtype, body))

    def _v2(self, r, buf, addr):
        flags = buf[addr + 5]
        o = addr + 6
        if flags & 0x20:
            o += 16
        if flags & 0x10:
            o += 4
        csz = 1 << (flags & 3)
        size = r.u(o, csz)
        o += csz
        order = 2 if flags & 0x04 else 0
        blocks = [(o, size)]
        while blocks:
            o, size = blocks.pop(0)
            end = o + size
            while o + 4 + order <= end:
                mtype, msize = buf[o], r.u(o + 1, 2)
                body = bytes(buf[o + 4 + order:o + 4 + order + msize])
                o += 4 + order + msize
                if mtype == 0x10:
                    caddr = r.addr_bytes(body, 0)
                    clen = int.from_bytes(body[r.osz:r.osz + r.lsz], 'little')
                    if buf[caddr:caddr + 4] != b'OCHK':
                        raise ValueError('bad object header continuation at %d' % caddr)
                    blocks.append((caddr + 4, clen - 8))      # minus signature and checksum
                elif mtype != 0:
                    self.msgs.append((mtype, body))

    def find(self, mtype):
        return [b for t, b in self.msgs if t == mtype]


class Dataset:
    """Read-only dataset: ``shape``, ``dtype``, ``ds[()]``, ``ds[lo:hi]`` (leading axis), ``len(ds)``."""

    def __init__(self, f, obj, name):
        self._f, self.name = f, name
        r = f._r
        self.shape = _parse_dataspace(r, obj.find(0x01)[0])
        self.dtype = _parse_datatype(obj.find(0x03)[0])
        lay = obj.find(0x08)[0]
        self._filters = _parse_filters(obj.find(0x0B)[0]) if obj.find(0x0B) else []
        ver = lay[0]
        if ver in (1, 2):
            # HDF5 1.6-era message: version, dimensionality, class, 5 reserved, [address], 4-byte dims
            # (+ element size for chunked), [compact: size + data]
            rank, cls = lay[1], lay[2]
            self._kind = cls
            o = 8
            if cls != 0:
                a = r.addr_bytes(lay, o)
                o += r.osz
            dims = tuple(int.from_bytes(lay[o + 4 * i:o + 4 * i + 4], 'little') for i in range(rank))
            o += 4 * rank
            if cls == 0:
                n = int.from_bytes(lay[o:o + 4], 'little')
                self._compact = lay[o + 4:o + 4 + n]
            elif cls == 1:
                self._addr = a
            elif cls == 2:
                self._btree, self._chunk = a, dims[:-1]
            else:
                raise NotImplementedError('data layout class %d' % cls)
            return
        cls = lay[1]
        if ver not in (3, 4):
            raise NotImplementedError('data layout message version %d' % ver)
        self._kind = cls
        if cls == 0:
            n = int.from_bytes(lay[2:4], 'little')
            self._compact = lay[4:4 + n]
        elif cls == 1:
            self._addr = r.addr_bytes(lay, 2)
        elif cls == 2:
            if ver == 4:
                raise NotImplementedError('version-4 chunked layout (libver="latest" chunk indexes) is not supported; '
                                          'BehaveNet writes contiguous datasets')
            rank = lay[2]
            self._btree = r.addr_bytes(lay, 3)
            o = 3 + r.osz
            self._chunk = tuple(int.from_bytes(lay[o + 4 * i:o + 4 * i + 4], 'little') for i in range(rank - 1))
        else:
            raise NotImplementedError('data layout class %d' % cls)

    def __len__(self):
        if not self.shape:
            raise TypeError('scalar dataset has no length')
        return self.shape[0]

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def _row_bytes(self):
        return int(np.prod(self.shape[1:], dtype=np.int64)) * self.dtype.itemsize

    def _read_rows(self, lo, hi):
        buf = self._f._r.buf
        shape = (hi - lo,) + self.shape[1:]
        if self._kind == 0:
            a = np.frombuffer(self._compact, self.dtype, self.size).reshape(self.shape)
            return a[lo:hi].copy()
        if self._kind == 1:
            if self._addr is None:          # never written: the fill value (zero)
                return np.zeros(shape, self.dtype)
            rb = self._row_bytes()
            start = self._addr + lo * rb
            return np.frombuffer(buf[start:start + (hi - lo) * rb], self.dtype).reshape(shape).copy()
        return self._read_chunked()[lo:hi]

    def _read_chunked(self):
        out = np.zeros(self.shape, self.dtype)
        if self._btree is None:
            return out
        r, buf = self._f._r, self._f._r.buf
        rank = len(self.shape)
        stack = [self._btree]
        while stack:
            node = stack.pop()
            if buf[node:node + 4] != b'TREE' or buf[node + 4] != 1:
                raise ValueError('bad chunk B-tree node at %d' % node)
            level, used = buf[node + 5], r.u(node + 6, 2)
            o = node + 8 + 2 * r.osz
            ksz = 8 + 8 * (rank + 1)
            for _ in range(used):
                nbytes, mask = r.u(o, 4), r.u(o + 4, 4)
                offs = [r.u(o + 8 + 8 * i, 8) for i in range(rank)]
                child = r.addr(o + ksz)
                o += ksz + r.osz
                if level > 0:
                    stack.append(child)
                    continue
                raw = bytes(buf[child:child + nbytes])
                for i, (fid, cd) in reversed(list(enumerate(self._filters))):
                    if mask & (1 << i):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else self.dtype.itemsize
                        raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                    elif fid == 3:
                        raw = raw[:-4]
                    else:
                        raise NotImplementedError('HDF5 filter id %d' % fid)
                chunk = np.frombuffer(raw, self.dtype).reshape(self._chunk)
                sel = tuple(slice(s, min(s + c, d)) for s, c, d in zip(offs, self._chunk, self.shape))
                out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out

    def __getitem__(self, key):
        if key == () or key is Ellipsis:
            if not self.shape:
                return self._read_scalar()
            return self._read_rows(0, self.shape[0])
        if isinstance(key, slice) and self.shape:
            lo, hi, step = key.indices(self.shape[0])
            if step == 1:
                return self._read_rows(lo, max(lo, hi))
        if not self.shape:
            raise IndexError('scalar dataset')
        return self._read_rows(0, self.shape[0])[key]

    def _read_scalar(self):
        if self._kind == 0:
            return np.frombuffer(self._compact, self.dtype, 1)[0]
        if self._kind == 1:
            buf = self._f._r.buf
            if self._addr is None:
                return self.dtype.type(0)
            return np.frombuffer(buf[self._addr:self._addr + self.dtype.itemsize], self.dtype)[0]
        return self._read_chunked()[()]

    def __array__(self, dtype=None, copy=None):
        a = self[()]
        return a if dtype is None else np.asarray(a, dtype)


class Group:
    """Read-only mapping of link names to groups / datasets."""

    def __init__(self, f, obj, name):
        self._f, self._obj, self.name = f, obj, name
        self._links = None

    def _load(self):
        if self._links is not None:
            return self._links
        r, buf = self._f._r, self._f._r.buf
        links = {}
        st = self._obj.find(0x11)
        if st:                                   # old style: symbol table = B-tree of symbol nodes + local heap
            btree, heap = r.addr_bytes(st[0], 0), r.addr_bytes(st[0], r.osz)
            if buf[heap:heap + 4] != b'HEAP':
                raise ValueError('bad local heap at %d' % heap)
            data = r.addr(heap + 8 + 2 * r.lsz)
            stack = [btree]
            while stack:
                node = stack.pop()
                if buf[node:node + 4] == b'SNOD':
                    n = r.u(node + 6, 2)
                    esz = 2 * r.osz + 24
                    for i in range(n):
                        e = node + 8 + i * esz
                        noff = r.u(e, r.osz)
                        end = buf.find(b'\0', data + noff)
                        links[bytes(buf[data + noff:end]).decode('utf-8')] = r.addr(e + r.osz)
                    continue
                if buf[node:node + 4] != b'TREE' or buf[node + 4] != 0:
                    raise ValueError('bad group B-tree node at %d' % node)
                used = r.u(node + 6, 2)
                o = node + 8 + 2 * r.osz + r.lsz            # skip key 0
                for _ in range(used):
                    stack.append(r.addr(o))
                    o += r.osz + r.lsz
        else:                                    # new style: link messages, or dense storage in a fractal heap
            for body in self._obj.find(0x06):
                name, addr, _ = _parse_link(r, body, 0)
                if addr is not None:
                    links[name] = addr
            for body in self._obj.find(0x02):
                flags = body[1]
                o = 2 + (8 if flags & 1 else 0)
                heap = r.addr_bytes(body, o)
                if heap is not None:
                    links.update(_fractal_heap_links(r, heap))
        self._links = links
        return links

    def keys(self):
        return sorted(self._load())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def __contains__(self, name):
        try:
            self[name]
        except KeyError:
            return False
        return True

    def __getitem__(self, path):
        node = self
        for part in [p for p in str(path).split('/') if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError("Unable to open object (object '%s' doesn't exist)" % part)
            node = node._f._open(links[part], node.name.rstrip('/') + '/' + part)
        return node


def _parse_link(r, b, o):
    """Link message (type 6) at b[o:] -> (name, object header address or None for soft/external links, next o)."""
    if b[o] != 1:
        raise ValueError('link message version %d' % b[o])
    flags = b[o + 1]
    o += 2
    ltype = 0
    if flags & 0x08:
        ltype = b[o]
        o += 1
    if flags & 0x04:
        o += 8
    if flags & 0x10:
        o += 1
    nsz = 1 << (flags & 3)
    nlen = int.from_bytes(b[o:o + nsz], 'little')
    o += nsz
    name = bytes(b[o:o + nlen]).decode('utf-8')
    o += nlen
    if ltype == 0:
        return name, r.addr_bytes(b, o), o + r.osz
    ln = int.from_bytes(b[o:o + 2], 'little')           # soft / external / user-defined: length-prefixed value
    return name, None, o + 2 + ln


def _fractal_heap_links(r, heap):
    """Every link message stored in the managed blocks of a fractal heap (dense link storage).  The blocks of a
    write-once file are filled front to back and zero-initialised, so the messages of a direct block are read
    consecutively until the version byte stops being 1."""
    buf = r.buf
    if buf[heap:heap + 4] != b'FRHP':
        raise ValueError('bad fractal heap header at %d' % heap)
    o = heap + 5
    o += 2                                              # heap ID length
    filt_len = r.u(o, 2)                                # encoded length of the I/O filter pipeline
    o += 2
    flags = buf[o]
    o += 1
    o += 4                                              # max size of managed objects
    o += r.lsz + r.osz                                  # next huge id, huge-object B-tree
    o += r.lsz + r.osz                                  # free space, free-space manager
    o += 4 * r.lsz                                      # managed space, allocated space, iterator offset, n objects
    o += 4 * r.lsz                                      # huge size / count, tiny size / count
    width = r.u(o, 2)
    o += 2
    start = r.length(o)                                 # starting block size
    o += r.lsz
    max_direct = r.length(o)
    o += r.lsz
    max_heap_bits = r.u(o, 2)
    o += 2
    o += 2                                              # starting rows of the root indirect block
    root = r.addr(o)
    o += r.osz
    root_rows = r.u(o, 2)
    if filt_len:
        raise NotImplementedError('filtered fractal heap')
    off_sz = (max_heap_bits + 7) // 8
    head = 5 + r.osz + off_sz + (4 if flags & 2 else 0)
    links = {}

    def direct(addr, size):
        if buf[addr:addr + 4] != b'FHDB':
            raise ValueError('bad fractal heap direct block at %d' % addr)
        b = bytes(buf[addr:addr + size])
        o = head
        while o + 4 < size and b[o] == 1:
            name, a, o = _parse_link(r, b, o)
            if a is not None:
                links[name] = a

    def row_size(row):
        return start << max(0, row - 1)

    def indirect(addr, nrows):
        if buf[addr:addr + 4] != b'FHIB':
            raise ValueError('bad fractal heap indirect block at %d' % addr)
        o = addr + 5 + r.osz + off_sz
        for row in range(nrows):
            size = row_size(row)
            for _ in range(width):
                child = r.addr(o)
                o += r.osz
                if child is None:
                    continue
                if size <= max_direct:
                    direct(child, size)
                else:                                   # rows of a child indirect block covering `size` bytes
                    indirect(child, (size // (start * width)).bit_length())

    if root is not None:
        if root_rows == 0:
            direct(root, start)
        else:
            indirect(root, root_rows)
    return links


class File(Group):
    """``hdf5_lite.File(path)`` -- read-only; extra h5py keywords (``libver``, ``swmr``) are accepted and ignored."""

    def __init__(self, path, mode='r', **kwargs):
        if mode != 'r':
            raise ValueError('hdf5_lite.File is read-only (use hdf5_lite.write for fixtures)')
        self.filename = str(path)
        self._fh = open(path, 'rb')
        try:
            self._map = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise OSError('Unable to open file (empty file): %s' % path)
        self.unpinned_format = False
        self._cache = {}
        buf = self._map
        base = 0
        while buf[base:base + 8] != _SIG:               # the superblock may follow a user block of 512 * 2^k bytes
            base = 512 if base == 0 else base * 2
            if base + 8 > len(buf):
                self.close()
                raise OSError('Unable to open file (file signature not found): %s' % path)
        # (the base address field always equals the superblock's own position: addresses are relative to it)
        r = self._r = _Reader(buf, base)
        ver = buf[base + 8]
        if ver in (0, 1):
            r.osz, r.lsz = buf[base + 13], buf[base + 14]
            o = base + 24 + (4 if ver == 1 else 0)
            root = r.addr(o + 4 * r.osz + r.osz)         # root symbol-table entry: name offset, header address
        elif ver in (2, 3):
            self.unpinned_format = True
            r.osz, r.lsz = buf[base + 9], buf[base + 10]
            root = r.addr(base + 12 + 3 * r.osz)
        else:
            self.close()
            raise NotImplementedError('HDF5 superblock version %d' % ver)
        Group.__init__(self, self, _Object(self, root), '/')

    def _open(self, addr, name):
        node = self._cache.get(addr)
        if node is None:
            obj = _Object(self, addr)
            node = Dataset(self, obj, name) if obj.find(0x01) and obj.find(0x08) else Group(self, obj, name)
            self._cache[addr] = node
        return node

    def close(self):
        m, self._map = getattr(self, '_map', None), None
        if m is not None:
            self._cache = {}
            self._r = None
            try:
                m.close()
            except BufferError:          # arrays are copies, so no exported views should exist
                pass
        if not self._fh.closed:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---- fixture writer (old-style files: the pinned format) ----------------------------------------------------------

def _datatype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in 'iu':
        bits0 = (0x08 if dt.kind == 'i' else 0)
        props = struct.pack('<HH', 0, 8 * dt.itemsize)
        cls = 0
    elif dt.kind == 'f' and dt.itemsize in (4, 8):
        # IEEE little-endian: padding 0, mantissa normalisation = implied msb (2 << 4), sign bit position in byte 2
        sign, ebits, mbits, bias = (31, 8, 23, 127) if dt.itemsize == 4 else (63, 11, 52, 1023)
        bits0 = 0x20
        props = struct.pack('<HHBBBBI', 0, 8 * dt.itemsize, mbits, ebits, 0, mbits, bias)
        cls = 1
        return bytes([0x10 | cls, bits0, sign, 0]) + struct.pack('<I', dt.itemsize) + props
    else:
        raise NotImplementedError('hdf5_lite.write: dtype %s' % dt)
    return bytes([0x10 | cls, bits0, 0, 0]) + struct.pack('<I', dt.itemsize) + props


def _msg(mtype, body, flags=0):
    pad = (-len(body)) % 8
    return struct.pack('<HHB3x', mtype, len(body) + pad, flags) + body + b'\0' * pad


def _object_header(msgs):
    body = b''.join(msgs)
    return struct.pack('<BxHII4x', 1, len(msgs), 1, len(body)) + body


def write(path, tree, userblock=0):
    """Write ``tree`` -- nested dicts whose leaves are numpy arrays (integers or float32/float64) -- as an
    old-style HDF5 file: superblock 0, one symbol-table node per group, contiguous little-endian datasets."""
    if userblock and (userblock < 512 or userblock & (userblock - 1)):
        raise ValueError('user block size must be a power of two >= 512')
    out = bytearray(96)                                   # superblock placeholder
    max_links = [1]

    def align():
        out.extend(b'\0' * ((-len(out)) % 8))

    def put(b):
        align()
        a = len(out)
        out.extend(b)
        return a

    def emit(node):
        if not isinstance(node, dict):
            a = np.asarray(node)
            a = a if a.flags.c_contiguous else a.copy()
            if a.dtype.byteorder == '>':
                a = a.astype(a.dtype.newbyteorder('<'))
            data = put(a.tobytes()) if a.size else _UNDEF
            space = struct.pack('<BBB5x', 1, a.ndim, 0) + b''.join(struct.pack('<Q', d) for d in a.shape)
            msgs = [_msg(0x01, space), _msg(0x03, _datatype_msg(a.dtype), 1),
                    _msg(0x05, struct.pack('<BBBB', 2, 2, 2, 0)),
                    _msg(0x08, struct.pack('<BBQQ', 3, 1, data, a.nbytes))]
            return put(_object_header(msgs))
        names = sorted(node, key=lambda s: s.encode('utf-8'))
        max_links[0] = max(max_links[0], len(names))
        children = [emit(node[k]) for k in names]
        heap_data = bytearray(8)                          # offset 0: the empty name
        offs = []
        for k in names:
            offs.append(len(heap_data))
            nb = k.encode('utf-8') + b'\0'
            heap_data.extend(nb + b'\0' * ((-len(nb)) % 8))
        data_addr = put(bytes(heap_data))
        heap = put(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), _UNDEF, data_addr))
        snod = bytearray(b'SNOD' + struct.pack('<BxH', 1, len(names)))
        for off, child in zip(offs, children):
            snod.extend(struct.pack('<QQII16x', off, child, 0, 0))
        snod_addr = put(bytes(snod))
        if names:
            tree_node = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, _UNDEF, _UNDEF) + struct.pack('<QQQ', 0, snod_addr, offs[-1])
        else:
            tree_node = b'TREE' + struct.pack('<BBHQQ', 0, 0, 0, _UNDEF, _UNDEF) + struct.pack('<Q', 0)
        btree = put(tree_node)
        return put(_object_header([_msg(0x11, struct.pack('<QQ', btree, heap))]))

    if not isinstance(tree, dict):
        raise TypeError('the root of the tree is a group (dict)')
    root = emit(tree)
    align()
    leaf_k = max(4, (max_links[0] + 1) // 2)
    if leaf_k > 0xFFFF:
        raise ValueError('too many links in one group for a single symbol-table node')
    sb = _SIG + struct.pack('<BBBxBBBxHHI', 0, 0, 0, 0, 8, 8, leaf_k, 16, 0)
    sb += struct.pack('<QQQQ', userblock, _UNDEF, len(out), _UNDEF)
    sb += struct.pack('<QQII16x', 0, root, 0, 0)          # root entry without cached scratch-pad data
    out[:96] = sb
    with open(path, 'wb') as fh:
        fh.write(b'\0' * userblock)
        fh.write(bytes(out))
