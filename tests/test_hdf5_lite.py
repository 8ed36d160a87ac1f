"""The dependency-free HDF5 reader behind ``HDF5Source`` (behavenet_b200/data/hdf5_lite.py).

* pinned against a file written by the HDF5 library itself: scipy ships one MATLAB v7.3 (= HDF5) file with its
  test data, whose content scipy's own tests state (``testdouble`` = linspace(0, 2 pi, 9) as a column);
* fixture writer -> reader round trips over the BehaveNet layout (``<signal>/trial_%04i``, reference
  data_generator.py:253-303; docs/source/data_structure.rst:59-79), user block, chunked / filtered datasets;
* the input pipeline on an HDF5 session against the same trials held in memory;
* the ``libver='latest'`` structures (superblock 2, version-2 object headers, compact and dense links) on files
  assembled here from the specification -- self-consistency only, that branch is unpinned (see the module header).
"""

import os
import pickle
import struct
import zlib

import numpy as np
import pytest

from behavenet_b200.data import hdf5_lite as h5
from behavenet_b200.data import ArraySource, HDF5Source, PrefetchSessionsGenerator


def _scipy_file():
    import scipy.io
    return os.path.join(os.path.dirname(scipy.io.__file__), 'matlab', 'tests', 'data', 'testhdf5_7.4_GLNX86.mat')


def test_reads_a_file_written_by_the_hdf5_library():
    path = _scipy_file()
    if not os.path.exists(path):
        pytest.skip('scipy test data not installed')
    with h5.File(path, 'r', libver='latest', swmr=True) as f:
        assert f.keys() == ['testdouble'] and len(f) == 1 and 'testdouble' in f and 'nope' not in f
        ds = f['testdouble']
        assert ds.shape == (9, 1) and ds.dtype == np.dtype('<f8') and len(ds) == 9
        np.testing.assert_array_equal(ds[()], np.linspace(0, 2 * np.pi, 9)[:, None])
        np.testing.assert_array_equal(ds[2:5], np.linspace(0, 2 * np.pi, 9)[2:5, None])
        assert f.unpinned_format is False
        with pytest.raises(KeyError):
            f['missing']


def _trials(n, rng, c=1, h=12, w=10, ragged=True):
    lens = [5 + (i % 4 if ragged else 0) for i in range(n)]
    return {
        'images': [rng.randint(0, 256, (t, c, h, w)).astype(np.uint8) for t in lens],
        'masks': [(rng.rand(t, c, h, w) > 0.2).astype(np.float32) for t in lens],
        'labels': [rng.randn(t, 4).astype(np.float32) for t in lens],
    }


def _write_session(path, trials, **kw):
    h5.write(path, {sig: {'trial_%04i' % i: a for i, a in enumerate(arrs)} for sig, arrs in trials.items()}, **kw)


@pytest.mark.parametrize('userblock', [0, 512, 2048])
def test_writer_reader_round_trip(tmp_path, userblock):
    rng = np.random.RandomState(0)
    tree = {
        'images': {'trial_%04i' % i: rng.randint(0, 256, (3 + i % 5, 2, 6, 4)).astype(np.uint8) for i in range(300)},
        'neural': {'trial_0000': rng.randn(7, 3).astype(np.float32), 'trial_0001': rng.randn(2, 3)},
        'ints': {'a': np.arange(-5, 5, dtype=np.int16), 'b': np.arange(6, dtype=np.int64).reshape(2, 3),
                 'c': np.array(7, dtype=np.uint32), 'empty': np.zeros((0, 3), np.float32)},
        'regions': {'indxs': {'region-0': np.arange(10), 'region-1': 10 + np.arange(15)}},
        'nothing': {},
        'top': np.float64(3.5) * np.ones((2, 2)),
    }
    path = str(tmp_path / 'data.hdf5')
    h5.write(path, tree, userblock=userblock)
    with h5.File(path) as f:
        assert f.keys() == sorted(tree)
        assert len(f['images']) == 300 and f['images'].keys()[:2] == ['trial_0000', 'trial_0001']

        def check(node, ref):
            if isinstance(ref, dict):
                assert isinstance(node, h5.Group) and node.keys() == sorted(ref)
                for k in ref:
                    check(node[k], ref[k])
            else:
                ref = np.asarray(ref)
                assert isinstance(node, h5.Dataset) and node.shape == ref.shape and node.dtype == ref.dtype
                np.testing.assert_array_equal(node[()], ref)
        check(f, tree)
        a = tree['images']['trial_0007']
        np.testing.assert_array_equal(f['images/trial_0007'][1:3], a[1:3])
        np.testing.assert_array_equal(f['images']['trial_0007'][2:], a[2:])
        np.testing.assert_array_equal(f['images']['trial_0007'][:, 1], a[:, 1])
        np.testing.assert_array_equal(np.asarray(f['regions']['indxs']['region-1']), 10 + np.arange(15))
        assert f['ints']['c'][()] == 7
    with pytest.raises(ValueError):
        h5.File(path, 'w')
    with pytest.raises(NotImplementedError):
        h5.write(path, {'s': np.array(['a', 'b'])})


def test_not_an_hdf5_file(tmp_path):
    p = tmp_path / 'junk.hdf5'
    p.write_bytes(b'x' * 5000)
    with pytest.raises(OSError):
        h5.File(str(p))


def _chunked_file(path, data, chunk, deflate, shuffle):
    """An old-style file with one chunked dataset 'x' (version-3 layout, version-1 B-tree chunk index, optional
    shuffle + deflate pipeline), assembled from the specification around hdf5_lite's own group writer."""
    h5.write(path, {'x': np.zeros(1, data.dtype)})
    raw = bytearray(open(path, 'rb').read())
    es = data.dtype.itemsize
    rank = data.ndim
    keys = []
    grid = [range(0, d, c) for d, c in zip(data.shape, chunk)]
    for idx in np.ndindex(*[len(g) for g in grid]):
        offs = [g[i] for g, i in zip(grid, idx)]
        block = np.zeros(chunk, data.dtype)
        sel = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, chunk, data.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sel)] = data[sel]
        b = block.tobytes()
        if shuffle:
            b = np.frombuffer(b, np.uint8).reshape(-1, es).T.tobytes()
        if deflate:
            b = zlib.compress(b, 4)
        while len(raw) % 8:
            raw.append(0)
        keys.append((len(b), offs, len(raw)))
        raw.extend(b)
    while len(raw) % 8:
        raw.append(0)
    btree = len(raw)
    node = b'TREE' + struct.pack('<BBHQQ', 1, 0, len(keys), h5._UNDEF, h5._UNDEF)
    for nbytes, offs, addr in keys:
        node += struct.pack('<II', nbytes, 0) + b''.join(struct.pack('<Q', o) for o in offs + [0]) + struct.pack('<Q', addr)
    node += struct.pack('<II', 0, 0) + b''.join(struct.pack('<Q', d) for d in list(data.shape) + [0])
    raw.extend(node)
    while len(raw) % 8:
        raw.append(0)
    filters = []
    if shuffle:
        filters.append(struct.pack('<HHHH', 2, 0, 0, 1) + struct.pack('<I', es) + b'\0' * 4)
    if deflate:
        filters.append(struct.pack('<HHHH', 1, 0, 0, 1) + struct.pack('<I', 4) + b'\0' * 4)
    space = struct.pack('<BBB5x', 1, rank, 0) + b''.join(struct.pack('<Q', d) for d in data.shape)
    layout = struct.pack('<BBB', 3, 2, rank + 1) + struct.pack('<Q', btree) + b''.join(
        struct.pack('<I', c) for c in list(chunk) + [es])
    msgs = [h5._msg(0x01, space), h5._msg(0x03, h5._datatype_msg(data.dtype), 1), h5._msg(0x08, layout)]
    if filters:
        msgs.append(h5._msg(0x0B, struct.pack('<BB6x', 1, len(filters)) + b''.join(filters)))
    hdr = len(raw)
    raw.extend(h5._object_header(msgs))
    # repoint the root group's only link at the new object header
    snod = raw.find(b'SNOD')
    raw[snod + 8 + 8:snod + 8 + 16] = struct.pack('<Q', hdr)
    open(path, 'wb').write(bytes(raw))


@pytest.mark.parametrize('deflate,shuffle', [(False, False), (True, False), (True, True)])
def test_chunked_and_filtered_datasets(tmp_path, deflate, shuffle):
    rng = np.random.RandomState(1)
    data = rng.randint(0, 50, (11, 2, 9, 7)).astype(np.uint16)
    path = str(tmp_path / 'chunked.hdf5')
    _chunked_file(path, data, (4, 1, 5, 7), deflate, shuffle)
    with h5.File(path) as f:
        ds = f['x']
        assert ds.shape == data.shape and ds.dtype == data.dtype
        np.testing.assert_array_equal(ds[()], data)
        np.testing.assert_array_equal(ds[3:9], data[3:9])


def test_hdf5_source_feeds_the_generator_like_arrays_in_memory(tmp_path):
    """Same epoch from an HDF5 session and from the same trials in memory (reference semantics: images / 255,
    everything float32; data_generator.py:250-285)."""
    rng = np.random.RandomState(3)
    trials = _trials(40, rng)
    path = str(tmp_path / 'data.hdf5')
    _write_session(path, trials)
    src = HDF5Source(path, ['images', 'masks', 'labels'], lab='lab', expt='e', animal='a', session='s', backend='lite')
    assert src.backend == 'lite' and src.n_trials == 40 and src.trial_length(3) == 8
    np.testing.assert_array_equal(src.load('images', 5, 1, 4), trials['images'][5][1:4])
    src2 = pickle.loads(pickle.dumps(src))
    np.testing.assert_array_equal(src2.load('labels', 7), trials['labels'][7])
    mem = ArraySource(trials, lab='lab', expt='e', animal='a', session='s')
    g1 = PrefetchSessionsGenerator([src], device='cpu', rng_seed=4, depth=3)
    g2 = PrefetchSessionsGenerator([mem], device='cpu', rng_seed=4, depth=3)
    try:
        assert g1.n_tot_batches == g2.n_tot_batches
        for split in ('train', 'val', 'test'):
            g1.reset_iterators(split)
            g2.reset_iterators(split)
            for _ in range(g1.n_tot_batches[split]):
                (a, da), (b, db) = g1.next_batch(split), g2.next_batch(split)
                assert da == db and a['batch_idx'] == b['batch_idx']
                for sig in ('images', 'masks', 'labels'):
                    assert a[sig].dtype == b[sig].dtype and a[sig].shape == b[sig].shape
                    np.testing.assert_array_equal(np.asarray(a[sig]), np.asarray(b[sig]))
                assert float(a['images'].max()) <= 1.0
    finally:
        g1.close()
        g2.close()
    with pytest.raises(KeyError):
        HDF5Source(path, ['images', 'neural'], backend='lite')
    with pytest.raises(ValueError):
        HDF5Source(path, ['images'], backend='pytables')


# ---- libver='latest' structures, assembled from the specification (self-consistency; unpinned) ---------------------

def _ohdr2(msgs):
    body = b''.join(struct.pack('<BHB', t, len(b), 0) + b for t, b in msgs)
    assert len(body) < 256
    return b'OHDR' + bytes([2, 0x00, len(body)]) + body + b'\0\0\0\0'          # flags 0: 1-byte chunk size; checksum


def _link(name, addr):
    nb = name.encode()
    return bytes([1, 0x00, len(nb)]) + nb + struct.pack('<Q', addr)            # hard link, 1-byte name length


def _latest_file(path, arrays, dense):
    """Superblock 2 + version-2 object headers; the root group holds its links either as link messages (compact) or
    in the root direct block of a fractal heap (dense storage)."""
    out = bytearray(48)
    addrs = {}
    for name, a in arrays.items():
        while len(out) % 8:
            out.append(0)
        data = len(out)
        out.extend(a.tobytes())
        space = bytes([2, a.ndim, 0, 1]) + b''.join(struct.pack('<Q', d) for d in a.shape)
        msgs = [(0x01, space), (0x03, h5._datatype_msg(a.dtype)), (0x08, struct.pack('<BBQQ', 3, 1, data, a.nbytes))]
        addrs[name] = len(out)
        out.extend(_ohdr2(msgs))
    if dense:
        while len(out) % 8:
            out.append(0)
        heap = len(out)
        start_size, off_bytes = 512, 4
        dblock = heap + 256
        hdr = b'FRHP' + bytes([0]) + struct.pack('<HHB', 7, 0, 0x02) + struct.pack('<I', 4096)
        hdr += struct.pack('<QQ', 0, h5._UNDEF) + struct.pack('<QQ', 0, h5._UNDEF)
        hdr += struct.pack('<QQQQ', start_size, start_size, start_size, len(arrays)) + struct.pack('<QQQQ', 0, 0, 0, 0)
        hdr += struct.pack('<HQQHH', 4, start_size, 65536, 32, 1) + struct.pack('<QH', dblock, 0) + b'\0\0\0\0'
        out.extend(hdr.ljust(256, b'\0'))
        blk = b'FHDB' + bytes([0]) + struct.pack('<Q', heap) + b'\0' * off_bytes + b'\0\0\0\0'
        blk += b''.join(_link(n, a) for n, a in addrs.items())
        out.extend(blk.ljust(start_size, b'\0'))
        root_msgs = [(0x02, bytes([0, 0]) + struct.pack('<QQ', heap, h5._UNDEF))]
    else:
        root_msgs = [(0x02, bytes([0, 0]) + struct.pack('<QQ', h5._UNDEF, h5._UNDEF))]
        root_msgs += [(0x06, _link(n, a)) for n, a in addrs.items()]
    while len(out) % 8:
        out.append(0)
    root = len(out)
    out.extend(_ohdr2(root_msgs))
    out[:48] = h5._SIG + bytes([2, 8, 8, 0]) + struct.pack('<QQQQ', 0, h5._UNDEF, len(out), root) + b'\0\0\0\0'
    open(path, 'wb').write(bytes(out))


@pytest.mark.parametrize('dense', [False, True])
def test_latest_format_structures_from_the_specification(tmp_path, dense):
    rng = np.random.RandomState(2)
    arrays = {'trial_%04i' % i: rng.randint(0, 256, (2 + i, 3)).astype(np.uint8) for i in range(5)}
    arrays['f'] = rng.randn(4).astype(np.float32)
    path = str(tmp_path / 'latest.hdf5')
    _latest_file(path, arrays, dense)
    with h5.File(path, 'r', libver='latest', swmr=True) as f:
        assert f.unpinned_format is True
        assert f.keys() == sorted(arrays)
        for k, a in arrays.items():
            assert f[k].shape == a.shape and f[k].dtype == a.dtype
            np.testing.assert_array_equal(f[k][()], a)


def test_round_trip_property():
    """Random trees (nesting, names, dtypes, ranks, empty and scalar datasets) survive write -> read."""
    import tempfile
    from hypothesis import given, settings, strategies as st, HealthCheck

    dtypes = st.sampled_from(['u1', 'i1', '<u2', '<i2', '<u4', '<i4', '<u8', '<i8', '<f4', '<f8', '>i4', '>f8'])
    names = st.text(alphabet='abcdefghijklmnopqrstuvwxyzABC_0123456789-. é', min_size=1, max_size=12).filter(
        lambda s: s not in ('.', '..'))

    @st.composite
    def arrays(draw):
        dt = np.dtype(draw(dtypes))
        shape = tuple(draw(st.lists(st.integers(0, 5), min_size=0, max_size=4)))
        n = int(np.prod(shape, dtype=np.int64))
        seed = draw(st.integers(0, 2 ** 31 - 1))
        rng = np.random.RandomState(seed)
        a = (rng.randn(n) * 100).astype(dt) if dt.kind == 'f' else rng.randint(0, 100, n).astype(dt)
        return a.reshape(shape)

    trees = st.recursive(arrays(), lambda kids: st.dictionaries(names, kids, max_size=6), max_leaves=12)

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.dictionaries(names, trees, max_size=6), st.sampled_from([0, 512]))
    def run(tree, userblock):
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, 't.hdf5')
            h5.write(path, tree, userblock=userblock)

            def check(node, ref):
                if isinstance(ref, dict):
                    assert sorted(node.keys()) == sorted(ref)
                    for k, v in ref.items():
                        check(node[k], v)
                else:
                    want = ref.dtype.newbyteorder('<') if ref.dtype.byteorder == '>' else ref.dtype      # written little-endian
                    assert node.shape == ref.shape and node.dtype == want
                    np.testing.assert_array_equal(node[()], ref)
            with h5.File(path) as f:
                check(f, tree)
    run()
