"""Input pipeline (behavenet_b200.data): the reference generator's protocol, its trial split against golden
outputs of the reference's own ``split_trials``, epoch coverage, frame sharding; on a GPU also the pinned /
side-stream prefetch path and the raw-uint8 hand-off to the encoder."""

import os

import numpy as np
import pytest
import torch

from behavenet_b200.data import ArraySource, PrefetchSessionsGenerator, split_trials
from tests.helpers import load_golden


def _session(n_trials, seed, c=1, h=8, w=8, name='s', ragged=True):
    rng = np.random.RandomState(seed)
    lens = [int(rng.randint(3, 9)) if ragged else 6 for _ in range(n_trials)]
    return ArraySource({
        'images': [rng.randint(0, 256, (T, c, h, w)).astype(np.uint8) for T in lens],
        'masks': [(rng.rand(T, c, h, w) > 0.2).astype(np.float32) for T in lens],
        'labels': [rng.randn(T, 3).astype(np.float64) for T in lens],
    }, lab='lab', expt='expt', animal='mouse', session=name)


def test_split_trials_matches_reference_golden():
    gold = load_golden('split_trials')
    for i, (n, seed, tr, va, te, gap) in enumerate(gold['cases']):
        got = split_trials(int(n), rng_seed=int(seed), train_tr=int(tr), val_tr=int(va), test_tr=int(te),
                           gap_tr=int(gap))
        for k in ('train', 'val', 'test'):
            assert np.array_equal(got[k], gold['%d_%s' % (i, k)]), (i, k)
    with pytest.raises(ValueError):
        split_trials(5)


def test_generator_protocol_and_epoch_coverage_cpu():
    srcs = [_session(40, 0, name='a'), _session(23, 1, name='b')]
    gen = PrefetchSessionsGenerator(srcs, device='cpu', rng_seed=3, depth=2)
    assert gen.n_datasets == 2 and len(gen) == 2
    assert gen.n_tot_batches == {'train': 32 + 16, 'val': 4 + 2, 'test': 4 + 2}
    np.testing.assert_allclose(gen.batch_ratios, [32 / 48, 16 / 48])
    assert gen.datasets[1].sess_str == 'lab_expt_mouse_b'
    for epoch in range(2):
        for dtype in ('train', 'val'):
            gen.reset_iterators(dtype)
            seen = [[], []]
            for _ in range(gen.n_tot_batches[dtype]):
                data, sess = gen.next_batch(dtype)
                idx = int(data['batch_idx'].item())
                seen[sess].append(idx)
                src = srcs[sess]
                x = data['images']
                assert x.dtype == torch.float32 and x.shape[0] == 1 and x.shape[1] == src.trial_length(idx)
                # reference scaling: float32(u8) / 255 on the host (data_generator.py:258-263)
                assert np.array_equal(x[0].numpy(), src.load('images', idx).astype('float32') / 255)
                assert np.array_equal(data['masks'][0].numpy(), src.load('masks', idx))
                assert data['labels'].dtype == torch.float32 and tuple(data['labels'].shape[::2]) == (1, 3)
            for s in range(2):
                assert sorted(seen[s]) == sorted(gen.datasets[s].batch_idxs[dtype].tolist())
            with pytest.raises(StopIteration):
                gen.next_batch(dtype)
    gen.close()


def test_generator_same_order_for_same_seed_and_numpy_mode():
    a = PrefetchSessionsGenerator([_session(30, 0), _session(20, 1)], device='cpu', rng_seed=5, as_numpy=True)
    b = PrefetchSessionsGenerator([_session(30, 0), _session(20, 1)], device='cpu', rng_seed=5, as_numpy=True)
    oa = [(s, d['batch_idx']) for d, s in (a.next_batch('train') for _ in range(a.n_tot_batches['train']))]
    ob = [(s, d['batch_idx']) for d, s in (b.next_batch('train') for _ in range(b.n_tot_batches['train']))]
    assert oa == ob and len(set(oa)) == len(oa)
    d, s = a.next_batch('test')
    assert isinstance(d['images'], np.ndarray) and d['images'].dtype == np.float32 and d['images'].shape[0] == 1
    a.close()
    b.close()


def test_frame_sharding_partitions_every_trial():
    from behavenet_b200 import parallel
    src = _session(20, 2)
    whole = PrefetchSessionsGenerator([src], device='cpu', rng_seed=1)
    parts = []
    for r in range(3):
        g = PrefetchSessionsGenerator([src], device='cpu', rng_seed=1, shard_frames=True)
        g._frame_range = (lambda T, r=r: parallel.shard_range(T, 3, r))
        parts.append(g)
    for _ in range(whole.n_tot_batches['train']):
        d, _s = whole.next_batch('train')
        T = d['images'].shape[1]
        pieces = [p.next_batch('train')[0] for p in parts]
        assert all(int(q['batch_idx']) == int(d['batch_idx']) for q in pieces)
        assert [q['shard'][1] for q in pieces] == [T] * 3
        assert torch.equal(torch.cat([q['images'] for q in pieces], 1), d['images'])
        assert [q['shard'][0] for q in pieces] == list(np.cumsum([0] + [q['images'].shape[1] for q in pieces[:-1]]))
    for g in parts + [whole]:
        g.close()


def test_worker_errors_surface_in_next_batch():
    src = _session(20, 0)
    src._data['masks'][int(split_trials(20)['val'][0])] = None
    gen = PrefetchSessionsGenerator([src], device='cpu')
    with pytest.raises(TypeError):
        for _ in range(gen.n_tot_batches['val']):
            gen.next_batch('val')
    gen.close()


def test_hdf5_source_backend_selection(tmp_path):
    """h5py when it is installed (like the reference), the dependency-free reader otherwise; asking for h5py
    explicitly where it is absent fails at construction, not at the first trial."""
    from behavenet_b200.data import HDF5Source, hdf5_lite
    path = str(tmp_path / 'data.hdf5')
    hdf5_lite.write(path, {'images': {'trial_0000': np.zeros((3, 1, 4, 4), np.uint8)}})
    try:
        import h5py  # noqa: F401
    except ImportError:
        assert HDF5Source(path, ['images']).backend == 'lite'
        with pytest.raises(ImportError):
            HDF5Source(path, ['images'], backend='h5py')
    else:
        assert HDF5Source(path, ['images']).backend == 'h5py'       # (reads a file hdf5_lite wrote)
        assert HDF5Source(path, ['images'], backend='lite').n_trials == 1


@pytest.mark.gpu
def test_prefetch_to_device_and_raw_uint8_into_encoder():
    """Pinned staging + side-stream copies deliver the reference's values on the device; raw uint8 batches go
    straight into the encoder's byte loader and give the same latents as the scaled float frames."""
    import copy
    from oracle import cae_oracle as co
    from behavenet_b200.models import AE
    src = _session(30, 4, c=1, h=64, w=48, ragged=True)
    gen = PrefetchSessionsGenerator([src], device='cuda', rng_seed=2, depth=3)
    raw = PrefetchSessionsGenerator([src], device='cuda', rng_seed=2, depth=3, raw_uint8=True)
    hp = co.make_hparams(1, 64, 48, 6)
    model = AE(copy.deepcopy(hp))
    model.load_state_dict(co.init_state_dict(hp, seed=0))
    model.cuda().eval()
    for epoch in range(2):
        gen.reset_iterators('train')
        raw.reset_iterators('train')
        for _ in range(gen.n_tot_batches['train']):
            d, _s = gen.next_batch('train')
            r, _s = raw.next_batch('train')
            idx = int(d['batch_idx'].item())
            assert int(r['batch_idx'].item()) == idx
            assert d['images'].is_cuda and d['images'].dtype == torch.float32
            assert r['images'].dtype == torch.uint8
            assert np.array_equal(d['images'][0].cpu().numpy(), src.load('images', idx).astype('float32') / 255)
            assert np.array_equal(d['masks'][0].cpu().numpy(), src.load('masks', idx))
            with torch.no_grad():
                zf = model.encoding(d['images'][0])[0]
                z8 = model.encoding(r['images'][0])[0]
            assert torch.equal(zf, z8)
    gen.close()
    raw.close()


def test_multi_session_generator_groups_distinct_sessions():
    """ConcatSessionsGeneratorMulti protocol (data_generator.py:636-800): training calls return one trial from
    each of n_sessions_per_batch different sessions, every trial at most once per pass, (None, None) when the
    sessions run out; val / test calls return single trials."""
    from behavenet_b200.data import PrefetchSessionsGeneratorMulti
    srcs = [_session(30, 0, name='a'), _session(20, 1, name='b'), _session(20, 2, name='c')]
    gen = PrefetchSessionsGeneratorMulti(srcs, n_sessions_per_batch=2, device='cpu', rng_seed=1)
    single_total = sum(len(d.batch_idxs['train']) for d in gen.datasets)
    assert gen.n_tot_batches['train'] == single_total // 2
    for epoch in range(2):
        gen.reset_iterators('train')
        seen = [set() for _ in srcs]
        n_groups = 0
        while True:
            samples, sessions = gen.next_batch('train')
            if samples is None:
                assert sessions is None
                break
            n_groups += 1
            assert len(samples) == 2 and len(set(sessions)) == 2
            for smp, s in zip(samples, sessions):
                idx = int(smp['batch_idx'])
                assert idx in set(gen.datasets[s].batch_idxs['train'].tolist()) and idx not in seen[s]
                seen[s].add(idx)
                assert np.array_equal(smp['images'][0].numpy(), srcs[s].load('images', idx).astype('float32') / 255)
        assert n_groups >= 16                      # the smaller two sessions (16 train trials each) bound the pass
        assert gen.next_batch('train') == (None, None)        # stays spent until reset_iterators
    d, s = gen.next_batch('val')
    assert isinstance(d, dict) and d['images'].shape[0] == 1
    d, s = gen.next_batch('train', return_multiple=False)
    assert isinstance(d, dict)
    with pytest.raises(NotImplementedError):
        PrefetchSessionsGeneratorMulti(srcs, n_sessions_per_batch=5, device='cpu')
    gen.close()


def test_export_states_and_latents_file_formats(tmp_path):
    """export_states (reference eval.py:120-181): every train / val / test trial decoded, gap trials left empty,
    the reference's pickle layout; with a batch-capable model one call per split, otherwise one per trial."""
    import pickle
    from behavenet_b200.fitting.eval import export_states
    rng = np.random.RandomState(0)
    lens = [int(rng.randint(4, 9)) for _ in range(24)]
    src = ArraySource({'ae_latents': [rng.randn(T, 3).astype(np.float32) for T in lens]},
                      lab='l', expt='e', animal='a', session='s')
    gen = PrefetchSessionsGenerator([src], device='cpu', rng_seed=0,
                                    trial_splits={'train_tr': 4, 'val_tr': 1, 'test_tr': 1, 'gap_tr': 1})

    class PerTrial:
        calls = 0

        def most_likely_states(self, y):
            PerTrial.calls += 1
            return (np.asarray(y)[:, 0] > 0).astype(np.int64)

    class Batched(PerTrial):
        batches = 0

        def most_likely_states_batch(self, ys):
            Batched.batches += 1
            return [(np.asarray(y)[:, 0] > 0).astype(np.int64) for y in ys]

    hp = {'model_class': 'arhmm', 'expt_dir': str(tmp_path), 'version': 0}
    for model, fname in ((PerTrial(), tmp_path / 'a.pkl'), (Batched(), tmp_path / 'b.pkl')):
        out = export_states(hp, gen, model, filename=str(fname))
        assert out == [str(fname)]
        with open(fname, 'rb') as f:
            d = pickle.load(f)
        assert set(d) == {'states', 'trials'} and len(d['states']) == 24
        used = np.concatenate([d['trials'][k] for k in ('train', 'val', 'test')])
        for i in range(24):
            if i in used:
                assert np.array_equal(d['states'][i], (src.load('ae_latents', i)[:, 0] > 0).astype(np.int64))
            else:
                assert d['states'][i].size == 0                      # gap trial
    assert PerTrial.calls == 12 and Batched.batches == 3       # 2 blocks of 4 + 1 + 1 trials (+ 3 gap trials each)
    (tmp_path / 'version_0').mkdir()
    out = export_states(hp, gen, Batched())
    assert out[0].endswith('version_0/l_e_a_s_states.pkl')
    gen.close()


def test_export_latents_branches_and_file_format(tmp_path):
    """export_latents (reference eval.py:6-118) with stand-in models on the CPU: the pickle layout, gap trials left
    empty, ``dataset=sess`` passed to the encoder, a conditional encoder fed frames + one-hot label images
    (eval.py:51-55, 70-71), AE-MSP latents exported through ``model.U`` (eval.py:79-81), PS-VAE subspaces
    concatenated (75-76), and the multi-session PS-VAE re-reading every trial through a single-session generator
    (vaes.py:1202-1273)."""
    import pickle
    import torch
    from behavenet_b200.fitting.eval import export_latents
    from behavenet_b200.data import PrefetchSessionsGeneratorMulti
    from behavenet_b200.models.vaes import MSPSVAE
    rng = np.random.RandomState(0)
    lens = [int(rng.randint(3, 7)) for _ in range(12)]

    def source(tag):
        return ArraySource({'images': [rng.rand(T, 1, 4, 4).astype(np.float32) for T in lens],
                            'labels_sc': [rng.rand(T, 2, 4, 4).astype(np.float32) for T in lens]},
                           lab='l', expt='e', animal='a', session=tag)
    src = source('s0')
    splits = {'train_tr': 2, 'val_tr': 1, 'test_tr': 1, 'gap_tr': 1}
    gen = PrefetchSessionsGenerator([src], device='cpu', rng_seed=0, trial_splits=splits)

    class Enc:
        def __init__(self, outs):
            self.outs, self.seen = outs, []

        def __call__(self, x, dataset=None):
            self.seen.append((tuple(x.shape[1:]), dataset))
            m = x.reshape(x.shape[0], -1).mean(1, keepdim=True)
            return tuple(m * (i + 1) for i in range(self.outs)) + ([], [])

    class Model:
        version = 0

        def __init__(self, mc, outs=1, **hp):
            self.hparams = dict(model_class=mc, expt_dir=str(tmp_path), n_ae_latents=outs, **hp)
            self.encoding = Enc(outs)

        def eval(self):
            return self

        def U(self, z):
            return -z

    def run(model, g=gen):
        f = tmp_path / ('%s.pkl' % model.hparams['model_class'])
        assert export_latents(g, model, filename=str(f)) == [str(f)]
        with open(f, 'rb') as fh:
            d = pickle.load(fh)
        assert set(d) == {'latents', 'trials'} and len(d['latents']) == 12
        return d

    d = run(Model('ae'))
    used = np.concatenate([d['trials'][k] for k in ('train', 'val', 'test')])
    assert 0 < len(used) < 12
    for i in range(12):
        if i in used:
            np.testing.assert_allclose(d['latents'][i][:, 0], src.load('images', i).reshape(lens[i], -1).mean(1), rtol=1e-6)
        else:
            assert d['latents'][i].size == 0
    m = Model('cond-ae', conditional_encoder=True)
    d = run(m)
    assert all(shape == (3, 4, 4) and ds == 0 for shape, ds in m.encoding.seen)      # 1 frame + 2 label channels
    i = int(used[0])
    both = np.concatenate([src.load('images', i), src.load('labels_sc', i)], 1)
    np.testing.assert_allclose(d['latents'][i][:, 0], both.reshape(lens[i], -1).mean(1), rtol=1e-6)
    d = run(Model('cond-ae-msp'))
    np.testing.assert_allclose(d['latents'][i][:, 0], -src.load('images', i).reshape(lens[i], -1).mean(1), rtol=1e-6)
    d = run(Model('ps-vae', outs=2))
    assert d['latents'][i].shape == (lens[i], 2)
    gen.close()
    # multi-session PS-VAE: the multi generator serves groups and skips trials; the export must not
    srcs = [source('s0'), source('s1')]
    multi = PrefetchSessionsGeneratorMulti(srcs, n_sessions_per_batch=2, device='cpu', rng_seed=0, trial_splits=splits)
    msps = Model('msps-vae', outs=3)
    msps.export_latents = lambda g, filename=None: MSPSVAE.export_latents(msps, g, filename=filename)
    (tmp_path / 'version_0').mkdir()
    out = export_latents(multi, msps)
    assert [os.path.basename(f) for f in out] == ['l_e_a_s0_latents.pkl', 'l_e_a_s1_latents.pkl']
    for f, sr in zip(out, srcs):
        with open(f, 'rb') as fh:
            d = pickle.load(fh)
        assert len(d['trials']['train']) == 12 and all(a.shape == (lens[j], 3) for j, a in enumerate(d['latents']))
        np.testing.assert_allclose(d['latents'][5][:, 1], 2 * sr.load('images', 5).reshape(lens[5], -1).mean(1), rtol=1e-6)
    multi.close()


def test_get_reconstruction_dispatch_per_model_class():
    """get_reconstruction (reference eval.py:284-376): which element of forward()'s tuple is the latent mean and
    which keyword arguments each model class receives; latent inputs go through the decoder only."""
    from behavenet_b200.fitting.eval import get_reconstruction

    class Mock:
        def __init__(self, mc, arity):
            self.hparams = {'model_class': mc, 'device': 'cpu'}
            self.arity, self.kwargs, self.decoded = arity, None, None

        def eval(self):
            return self

        def __call__(self, x, dataset=None, **kw):
            self.kwargs = kw
            return tuple(torch.full((x.shape[0], 2), float(i)) if i else x * 0 + 7 for i in range(self.arity))

        def decoding(self, z, pool_idx, outsize, dataset=None):
            self.decoded = z
            return torch.zeros(z.shape[0], 1, 4, 4)

        def get_inverse_transformed_latents(self, z, as_numpy=True):
            return z + 100

    x = torch.rand(3, 1, 4, 4)
    lab, lab2d = torch.ones(3, 2), torch.ones(3, 1, 4, 4)
    for mc, arity, pos, keys in [('ae', 2, 1, set()), ('cond-ae-msp', 3, 1, set()), ('vae', 4, 1, {'use_mean'}),
                                 ('beta-tcvae', 4, 1, {'use_mean'}), ('ps-vae', 5, 2, {'use_mean'}),
                                 ('msps-vae', 5, 2, {'use_mean'}), ('cond-ae', 2, 1, {'labels', 'labels_2d'}),
                                 ('cond-vae', 4, 1, {'labels', 'labels_2d'})]:
        m = Mock(mc, arity)
        ims, lat = get_reconstruction(m, x, return_latents=True, labels=lab, labels_2d=lab2d)
        assert set(m.kwargs) == keys and np.all(ims == 7) and np.all(lat == pos), mc
        assert isinstance(get_reconstruction(m, x.numpy()), np.ndarray)
    with pytest.raises(ValueError):
        get_reconstruction(Mock('nope', 2), x)
    z = torch.zeros(3, 2)
    m = Mock('ps-vae', 5)
    _, lat = get_reconstruction(m, z, return_latents=True)
    assert np.all(lat == 100) and torch.all(m.decoded == 100)
    _, lat = get_reconstruction(m, z, return_latents=True, apply_inverse_transform=False)
    assert np.all(lat == 0)
    m = Mock('cond-ae', 2)
    _, lat = get_reconstruction(m, z, return_latents=True, labels=lab)
    assert lat.shape == (3, 4) and np.all(lat[:, 2:] == 1)
    m = Mock('ae', 2)
    assert get_reconstruction(m, z).shape == (3, 1, 4, 4) and torch.all(m.decoded == 0)


def test_encode_trials_groups_and_orders_on_a_mock_model():
    """Host logic of fitting.eval.encode_trials on the CPU with a stand-in model: trials are grouped into launches of
    at most ``frames_per_launch`` frames (a single longer trial still goes alone), uint8 groups are scaled when the
    first layer cannot read bytes, mixed uint8 / float groups are unified, and the latents come back in trial order."""
    import torch
    from behavenet_b200.fitting.eval import encode_trials

    class Mock(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.hparams = {'model_class': 'ae', 'model_type': 'linear', 'n_ae_latents': 3}
            self.calls = []

        def encoding(self, x, dataset=None):
            assert x.dtype == torch.float32          # (no byte loader for this model: uint8 arrives scaled)
            self.calls.append(int(x.shape[0]))
            m = x.reshape(x.shape[0], -1).mean(1, keepdim=True)
            return torch.cat([m, 2 * m, m * 0 + x.shape[0]], 1), None, None

    rng = np.random.RandomState(0)
    lens = [5, 7, 40, 3, 3, 9]
    trials = [rng.randint(0, 256, (t, 1, 4, 4)).astype(np.uint8) for t in lens]
    trials[4] = trials[4].astype(np.float32) / 255            # a float trial next to uint8 ones
    model = Mock()
    model.train()
    lat, lengths = encode_trials(model, trials, frames_per_launch=16, device='cpu')
    assert lengths == lens and lat.shape == (sum(lens), 3) and model.training
    assert model.calls == [12, 40, 15]                       # [5, 7] | [40] | [3, 3, 9]
    want = np.concatenate([t.reshape(t.shape[0], -1).astype(np.float64).mean(1) / (255 if t.dtype == np.uint8 else 1)
                           for t in trials])
    np.testing.assert_allclose(lat[:, 0].numpy(), want, rtol=1e-6)
    np.testing.assert_allclose(lat[:, 1].numpy(), 2 * want, rtol=1e-6)
    assert encode_trials(model, [], device='cpu')[0].shape == (0, 3)
