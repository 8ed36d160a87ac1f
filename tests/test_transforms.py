"""behavenet_b200.data.transforms against the known-answer cases of the reference's own
tests/test_data/test_transforms.py (re-stated here; /root/reference is not read at run time)."""

import numpy as np
import pytest

from behavenet_b200.data import transforms


def _runs(sample):
    change = np.flatnonzero(np.diff(sample) != 0) + 1
    bounds = np.concatenate([[0], change, [len(sample)]])
    out = {}
    for a, b in zip(bounds[:-1], bounds[1:]):
        out.setdefault(sample[a], []).append(b - a)
    return {k: sorted(v) for k, v in out.items()}


def test_compose_select_then_zscore():
    t = transforms.Compose([transforms.SelectIdxs(np.array([0, 3])), transforms.ZScore()])
    s = t(np.random.RandomState(0).randn(100, 4))
    assert s.shape == (100, 2)
    assert np.allclose(s.mean(0), 0, atol=1e-3) and np.allclose(s.std(0), 1, atol=1e-3)
    assert 'SelectIndxs' in repr(t) and 'ZScore()' in repr(t)


def test_blockshuffle_keeps_run_statistics_and_is_seeded():
    t = transforms.BlockShuffle(0)
    signal = np.array([0, 0, 0, 1, 1, 1, 2, 2, 0, 0, 1, 1])
    s = t(signal)
    assert not np.all(signal == s)
    assert np.array_equal(np.bincount(signal), np.bincount(s))
    assert _runs(signal) == _runs(s)
    assert np.array_equal(s, t(signal))                              # re-seeded on every call
    assert np.all(np.isnan(t(np.array([0.0, np.nan, 1.0]))))
    # the reference's draw for this seed (one np.random.permutation of the 5 runs after np.random.seed(0))
    np.random.seed(0)
    order = np.random.permutation(5)
    runs = [np.arange(0, 3), np.arange(3, 6), np.arange(6, 8), np.arange(8, 10), np.arange(10, 12)]
    assert np.array_equal(s, signal[np.concatenate([runs[i] for i in order])])


def test_clipnormalize():
    for bad in (0, -2.3):
        with pytest.raises(ValueError):
            transforms.ClipNormalize(bad)
    t = transforms.ClipNormalize(2)
    signal = np.random.RandomState(1).randn(3, 3)
    assert np.max(t(signal)) <= 1
    signal[0, 0] = 3
    assert np.max(t(signal)) == 1


def test_makeonehot():
    t = transforms.MakeOneHot()
    already = np.array([[0, 1, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1]])
    assert np.all(t(already) == already)
    s = t(np.array([3, 3, 2, 2, 0]))
    assert np.all(s == np.array([[0, 0, 0, 1], [0, 0, 0, 1], [0, 0, 1, 0], [0, 0, 1, 0], [1, 0, 0, 0]]))
    assert np.all(np.isnan(t(np.array([1.0, np.nan, 0.0]))))


def test_makeonehot2d_reference_cases():
    t = transforms.MakeOneHot2D(4, 4)
    sp = np.zeros((3, 2, 4, 4))
    for idx in [(0, 0, 0, 1), (0, 1, 3, 2), (1, 0, 1, 0), (1, 1, 1, 2), (2, 0, 1, 3), (2, 1, 2, 0)]:
        sp[idx] = 1
    assert np.all(t(np.array([[1, 2, 0, 3], [0, 2, 1, 1], [3, 0, 1, 2]])) == sp)
    assert np.all(t(np.array([[1.2, 2.1, 0.1, 2.9], [0.2, 1.7, 1.1, 0.9], [3.2, 0.4, 1.3, 1.6]])) == sp)
    t3 = transforms.MakeOneHot2D(3, 3)                               # clipping at the borders
    sp = np.zeros((3, 2, 3, 3))
    for idx in [(0, 0, 0, 1), (0, 1, 2, 2), (1, 0, 1, 0), (1, 1, 1, 2), (2, 0, 1, 2), (2, 1, 2, 0)]:
        sp[idx] = 1
    assert np.all(t3(np.array([[1, 2, 0, 3], [-1, 2, 1, 1], [3, -2, 1, 4]])) == sp)
    sp = np.zeros((3, 2, 4, 4))                                      # NaNs go to pixel 0
    for idx in [(0, 0, 0, 1), (0, 1, 0, 2), (1, 0, 1, 0), (1, 1, 1, 2), (2, 0, 0, 3), (2, 1, 2, 0)]:
        sp[idx] = 1
    sig = np.array([[1, 2, 0, np.nan], [0, 2, 1, 1], [3, 0, np.nan, 2]])
    keep = sig.copy()
    assert np.all(t(sig) == sp)
    assert np.array_equal(sig, keep, equal_nan=True)                 # input untouched
    assert repr(t) == 'MakeOneHot2D(y_pixels=4, x_pixels=4)'


def test_motionenergy_selectidxs_threshold_zscore():
    rng = np.random.RandomState(2)
    signal = rng.randn(100, 4)
    me = transforms.MotionEnergy()(signal)
    assert me.shape == (100, 4) and np.all(me >= 0) and np.all(me[0] == 0)
    assert np.allclose(me[1:], np.abs(np.diff(signal, axis=0)))
    idxs = np.array([0, 3])
    assert np.all(transforms.SelectIdxs(idxs)(signal) == signal[:, idxs])
    with pytest.raises(ValueError):
        transforms.Threshold(1, 0)
    with pytest.raises(ValueError):
        transforms.Threshold(-1, 1)
    assert transforms.Threshold(0, 1)(rng.uniform(0, 4, (5, 4))).shape == (5, 4)
    sig = rng.uniform(2, 4, (5, 4))
    sig[:, 0] = 0
    out = transforms.Threshold(1, 1e3)(sig)
    assert out.shape == (5, 3) and out.dtype == np.float64 and np.array_equal(out, sig[:, 1:])
    z = transforms.ZScore()(10 + 0.3 * rng.randn(100, 3))
    assert np.allclose(z.mean(0), 0, atol=1e-3) and np.allclose(z.std(0), 1, atol=1e-3)


def test_generator_applies_transforms_for_conditional_encoder_inputs():
    """labels -> MakeOneHot2D frames ('labels_sc') inside the prefetch worker."""
    from behavenet_b200.data import ArraySource, PrefetchSessionsGenerator
    rng = np.random.RandomState(3)
    lens = [4, 6, 5, 7, 4, 5, 6, 4, 5, 6]
    labels = [rng.uniform(0, 8, (T, 4)) for T in lens]
    src = ArraySource({'images': [rng.randint(0, 256, (T, 1, 8, 8)).astype(np.uint8) for T in lens],
                       'labels': labels, 'labels_sc': [a.copy() for a in labels]})
    gen = PrefetchSessionsGenerator([src], device='cpu', transforms={'labels_sc': transforms.MakeOneHot2D(8, 8)})
    d, _ = gen.next_batch('train')
    idx = int(d['batch_idx'])
    assert tuple(d['labels_sc'].shape) == (1, lens[idx], 2, 8, 8)
    assert float(d['labels_sc'].sum()) == 2 * lens[idx]
    assert np.array_equal(d['labels_sc'][0].numpy(), transforms.MakeOneHot2D(8, 8)(labels[idx]).astype('float32'))
    gen.close()
