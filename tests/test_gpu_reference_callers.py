"""The reference's OWN callers, unmodified, driving the swapped-in classes ("must run
unchanged against the new model classes").

* ``behavenet.fitting.training.fit`` (reference training.py:244-461) -- optimizer set-up, epoch-0 evaluation
  without a step, per-epoch validation, ``copy.deepcopy`` of the best model around ``hparams = None``,
  ``model.save``, the test pass and the reference's ``export_latents`` -- runs with
  ``behavenet_b200.models.AE`` / ``PSVAE`` and ``behavenet_b200.data.PrefetchSessionsGenerator``; only the
  test-tube ``exp`` object is a stub.
* the model-construction / EM / evaluation / pickling body of ``arhmm_grid_search.main`` (reference
  arhmm_grid_search.py:131-209) is read from the reference file and executed verbatim with
  ``behavenet_b200.ssm`` bound to the name ``ssm``.

The reference package comes from baseline/_ref (installed by baseline/install_ref.sh; travels with gpurun)
or, in the authoring container, from /root/reference.
"""

import copy
import os
import pickle
import sys
import textwrap
import types

import numpy as np
import pytest
import torch

from oracle import cae_oracle as co
from oracle import arhmm_oracle as ao
from tests.helpers import ROOT

pytestmark = pytest.mark.gpu


def _reference_root():
    for root in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(root, 'behavenet', 'fitting')):
            return root
    pytest.skip('the reference package is not installed (run baseline/install_ref.sh)')


def _import_reference_training():
    root = _reference_root()
    sys.modules.setdefault('commentjson', types.ModuleType('commentjson'))
    if root not in sys.path:
        sys.path.insert(0, root)
    import behavenet.fitting.training as training
    assert os.path.realpath(training.__file__).startswith(os.path.realpath(root))
    return training


class StubExperiment:
    """What fit() uses of test_tube.Experiment: .version, .log(dict), .save()."""

    def __init__(self, version=0):
        self.version, self.rows, self.saves = version, [], 0

    def log(self, row):
        self.rows.append(dict(row))

    def save(self):
        self.saves += 1


def _structured_frames(n_trials, c, h, w, seed):
    rng = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing='ij')
    trials = []
    for _ in range(n_trials):
        T = int(rng.randint(10, 20))
        cx, cy = rng.rand(2)
        t = np.linspace(0, 1, T)[:, None, None, None]
        img = 0.5 + 0.4 * np.exp(-(((xx - cx - 0.2 * t) ** 2 + (yy - cy) ** 2) / 0.02))
        trials.append(np.clip(np.broadcast_to(img, (T, c, h, w)) * 255, 0, 255).astype(np.uint8))
    return trials


@pytest.mark.parametrize('model_class', ['ae', 'ps-vae'])
def test_reference_fit_runs_unchanged_with_the_b200_classes(model_class, tmp_path):
    from behavenet_b200.data import ArraySource, PrefetchSessionsGenerator
    from behavenet_b200.models import AE, PSVAE
    training = _import_reference_training()
    n_labels = 2 if model_class == 'ps-vae' else 0
    c = 2 if model_class == 'ps-vae' else 1
    hp = co.make_hparams(c, 32, 32, 6, model_class, n_labels)
    expt_dir = str(tmp_path)
    os.makedirs(os.path.join(expt_dir, 'version_0'))
    hp.update({'learning_rate': 1e-3, 'l2_reg': 0.0, 'enable_early_stop': True, 'early_stop_history': 10,
               'min_n_epochs': 1, 'max_n_epochs': 3, 'val_check_interval': 1, 'rng_seed_train': 0,
               'expt_dir': expt_dir, 'export_latents': True, 'save_last_model': True, 'device': 'cuda',
               'ps_vae.anneal_epochs': 2})
    frames = _structured_frames(20, c, 32, 32, seed=1)
    signals = {'images': frames}
    if n_labels:
        rng = np.random.RandomState(2)
        signals['labels'] = [rng.randn(f.shape[0], n_labels).astype(np.float32) for f in frames]
    src = ArraySource(signals, lab='lab', expt='expt', animal='mouse', session='s0')
    gen = PrefetchSessionsGenerator([src], device='cuda', rng_seed=0,
                                    trial_splits={'train_tr': 6, 'val_tr': 2, 'test_tr': 2, 'gap_tr': 0})
    torch.manual_seed(0)
    np.random.seed(0)
    model = (PSVAE if model_class == 'ps-vae' else AE)(hp)
    model.to('cuda')
    model.version = 0
    exp = StubExperiment(0)
    training.fit(hp, model, gen, exp, method='ae')                 # the reference's function, unmodified
    gen.close()
    train_rows = [r for r in exp.rows if r.get('dataset') == -1 and 'tr_loss' in r]
    val_rows = [r for r in exp.rows if r.get('dataset') == -1 and 'val_loss' in r]
    assert len(train_rows) == 4 and len(val_rows) == 4             # epochs 0..3
    assert all(np.isfinite(r['tr_loss']) for r in train_rows)
    assert train_rows[-1]['tr_loss'] < train_rows[0]['tr_loss']     # epoch 0 evaluates, epochs 1-3 train
    assert [r for r in exp.rows if 'test_loss' in r]
    vdir = os.path.join(expt_dir, 'version_0')
    for f in ('best_val_model.pt', 'last_model.pt', 'lab_expt_mouse_s0_latents.pkl'):
        assert os.path.exists(os.path.join(vdir, f)), f
    sd = torch.load(os.path.join(vdir, 'best_val_model.pt'), map_location='cpu')
    assert set(sd) == set(model.state_dict())
    with open(os.path.join(vdir, 'lab_expt_mouse_s0_latents.pkl'), 'rb') as fh:
        d = pickle.load(fh)
    assert len(d['latents']) == 20 and all(a.shape == (frames[i].shape[0], 6) for i, a in enumerate(d['latents']))
    assert model.hparams is hp                                      # restored after the deepcopy dance


def test_reference_arhmm_grid_search_body_runs_verbatim_with_the_b200_hmm(tmp_path):
    import behavenet_b200.ssm as ssm_b200
    root = _reference_root()
    path = os.path.join(root, 'behavenet', 'fitting', 'arhmm_grid_search.py')
    lines = open(path).read().split('\n')
    start = next(i for i, l in enumerate(lines) if l.strip() == "print('constructing model...', end='')")
    stop = next(i for i, l in enumerate(lines) if l.strip() == 'pickle.dump(hmm, f)')
    body = textwrap.dedent('\n'.join(lines[start:stop + 1]))
    assert 'ssm.HMM(' in body and "hmm.fit(latents['train'], method='em', num_iters=1, initialize=False)" in body
    assert 'hmm.most_likely_states(x)' in body and 'hmm.permute(perm)' in body

    K, D, lags = 4, 6, 1
    p = ao.synth_params(K, D, lags, seed=5, mix=0.1)
    rng = np.random.RandomState(3)

    def draw(n):
        return [ao.sample(p, int(rng.randint(80, 140)), rng)[1].astype(np.float32) for _ in range(n)]
    latents = {'train': draw(8), 'val': draw(3), 'test': draw(3)}
    exp = StubExperiment(0)
    os.makedirs(os.path.join(str(tmp_path), 'version_0'))
    hparams = {'rng_seed_model': 0, 'n_arhmm_states': K, 'n_iters': 3, 'arhmm_es_tol': 0, 'expt_dir': str(tmp_path)}
    ns = {
        'np': np, 'os': os, 'pickle': pickle, 'ssm': ssm_b200, 'hparams': hparams, 'exp': exp,
        'obs_dim': D, 'obs_type': 'ar', 'obs_kwargs': {'lags': lags}, 'obs_init_kwargs': {'localize': True},
        'transitions': 'sticky', 'transition_kwargs': {'kappa': 10.0},
        'latents': latents, 'latents_sess': {0: latents}, 'n_datasets': 1,
        'trial_idxs_sess': {0: {k: list(range(len(v))) for k, v in latents.items()}},
        'data_generator': types.SimpleNamespace(n_datasets=1),
        'export_hparams': lambda hp, e: None,
    }
    exec(compile(body, path, 'exec'), ns)                           # reference lines 131-209, verbatim
    hmm = ns['hmm']
    rows = [r for r in exp.rows if r.get('dataset') == -1]
    assert [r['epoch'] for r in rows] == [0, 1, 2, 3]
    tr = [r['tr_loss'] for r in rows]
    assert all(np.isfinite(tr)) and all(b <= a + 1e-6 * abs(a) for a, b in zip(tr[:-1], tr[1:]))   # EM is monotone
    assert len([r for r in exp.rows if 'test_loss' in r]) == 3
    assert hparams['training_completed'] is False and hmm.hparams is hparams
    with open(os.path.join(str(tmp_path), 'version_0', 'best_val_model.pt'), 'rb') as fh:
        clone = pickle.load(fh)
    zs = [hmm.most_likely_states(x) for x in latents['train']]
    usage = np.bincount(np.concatenate(zs), minlength=K)
    assert np.array_equal(usage, np.sort(usage)[::-1])              # states re-ordered by usage (lines 201-204)
    assert np.array_equal(clone.most_likely_states(latents['train'][0]), zs[0])


def test_reference_fit_with_the_linear_ae_on_an_hdf5_session(tmp_path):
    """The same unmodified fit() on the linear autoencoder (model_type='linear') fed from a ``data.hdf5`` in the
    reference's layout (``images/trial_%04i``, uint8) through HDF5Source's dependency-free reader."""
    from behavenet_b200.data import HDF5Source, PrefetchSessionsGenerator, hdf5_lite
    from behavenet_b200.models import AE
    training = _import_reference_training()
    hp = co.make_linear_hparams(1, 32, 32, 6)
    expt_dir = str(tmp_path)
    os.makedirs(os.path.join(expt_dir, 'version_0'))
    hp.update({'learning_rate': 1e-3, 'l2_reg': 0.0, 'enable_early_stop': True, 'early_stop_history': 10,
               'min_n_epochs': 1, 'max_n_epochs': 3, 'val_check_interval': 1, 'rng_seed_train': 0,
               'expt_dir': expt_dir, 'export_latents': True, 'save_last_model': True, 'device': 'cuda'})
    frames = _structured_frames(20, 1, 32, 32, seed=1)
    path = os.path.join(expt_dir, 'data.hdf5')
    hdf5_lite.write(path, {'images': {'trial_%04i' % i: f for i, f in enumerate(frames)}})
    src = HDF5Source(path, ['images'], lab='lab', expt='expt', animal='mouse', session='s0', backend='lite')
    gen = PrefetchSessionsGenerator([src], device='cuda', rng_seed=0,
                                    trial_splits={'train_tr': 6, 'val_tr': 2, 'test_tr': 2, 'gap_tr': 0})
    torch.manual_seed(0)
    np.random.seed(0)
    model = AE(hp)
    model.to('cuda')
    model.version = 0
    exp = StubExperiment(0)
    training.fit(hp, model, gen, exp, method='ae')
    gen.close()
    train_rows = [r for r in exp.rows if r.get('dataset') == -1 and 'tr_loss' in r]
    assert len(train_rows) == 4 and all(np.isfinite(r['tr_loss']) for r in train_rows)
    assert train_rows[-1]['tr_loss'] < train_rows[0]['tr_loss']
    vdir = os.path.join(expt_dir, 'version_0')
    sd = torch.load(os.path.join(vdir, 'best_val_model.pt'), map_location='cpu')
    assert set(sd) == set(model.state_dict())
    with open(os.path.join(vdir, 'lab_expt_mouse_s0_latents.pkl'), 'rb') as fh:
        d = pickle.load(fh)
    full = [z for z in d['latents'] if len(z)]
    assert len(full) == 20 and all(z.shape == (f.shape[0], 6) for z, f in zip(d['latents'], frames))
