"""CPU tests of the host-side mirror of the reference interfaces (no kernels are launched):
model construction / state_dict protocol, unsupported-variant errors, HMM host logic (M-step,
permute, pickling), sharding arithmetic, and the world-size-2 gloo path of the collectives."""

import copy
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import arhmm_oracle as ao
from oracle import cae_oracle as co
from tests.helpers import load_golden


def test_ae_state_dict_protocol_matches_reference_names():
    from behavenet_b200.models import AE
    hp = co.make_hparams(1, 128, 128, 12)
    model = AE(copy.deepcopy(hp))
    gold = load_golden('ae_128x128x1_l12_b3')
    ref_names = sorted({k[5:].split('#')[0] for k in gold if k.startswith('grad.')})
    assert sorted(model.state_dict().keys()) == ref_names
    sd = co.init_state_dict(hp)
    model.load_state_dict(sd)                      # shapes identical to the reference's
    assert model.hparams['hidden_layer_size'] == 12
    assert sum(p.numel() for p in model.parameters()) == 8758285     # SURVEY.md probe
    assert len(list(model.get_parameters())) == 24
    assert 'Encoder architecture' in str(model) and 'Decoder architecture' in str(model)
    assert model.decoding.conv_t_pads['convtranspose0'] is None
    assert model.decoding.conv_t_pads['convtranspose1'] == [1, 2, 1, 2]
    copy.deepcopy(model)
    pickle.loads(pickle.dumps(model))


def test_psvae_protocol():
    from behavenet_b200.models import PSVAE
    np.random.seed(0)
    hp = co.make_hparams(2, 32, 32, 8, 'ps-vae', 3)
    model = PSVAE(copy.deepcopy(hp))
    assert model.hparams['variational'] is True
    names = set(model.state_dict().keys())
    for k in ['encoding.A.weight', 'encoding.B.weight', 'encoding.D.weight', 'encoding.D.bias',
              'encoding.logvar.weight', 'encoding.FF.weight', 'decoding.FF.weight']:
        assert k in names
    assert not model.encoding.A.weight.requires_grad and not model.encoding.B.weight.requires_grad
    m = torch.cat([model.encoding.A.weight, model.encoding.B.weight], 0)
    assert torch.allclose(m @ m.T, torch.eye(8), atol=1e-5)           # orthogonal, frozen
    assert model.beta_vals.shape == (hp['max_n_epochs'] + 1,)
    hp2 = co.make_hparams(2, 32, 32, 2, 'ps-vae', 3)
    with pytest.raises(ValueError):
        PSVAE(hp2)
    hp3 = dict(co.make_hparams(2, 32, 32, 8, 'ps-vae', 3))
    hp3['ps_vae.anneal_epochs'] = 4
    model3 = PSVAE(hp3)
    assert model3.beta_vals[0] == 0 and model3.kl_anneal_vals[3] == 1


def test_vae_family_protocol():
    """VAE / BetaTCVAE keep the reference's constructor side effects, state_dict names, annealing
    tables (vaes.py:52-100, 380-409) and are deep-copyable / picklable like the AE."""
    from behavenet_b200.models import VAE, BetaTCVAE
    for mc, cls in (('vae', VAE), ('beta-tcvae', BetaTCVAE)):
        hp = co.make_hparams(1, 64, 48, 6, mc)
        model = cls(copy.deepcopy(hp))
        assert model.hparams['variational'] is True
        assert set(model.state_dict().keys()) == set(co.init_state_dict(hp, seed=0).keys())
        assert model.beta_vals.shape == (hp['max_n_epochs'] + 1,)
        copy.deepcopy(model)
        pickle.loads(pickle.dumps(model))
        with pytest.raises(RuntimeError):           # no CPU path
            model.loss({'images': torch.rand(1, 2, 1, 64, 48)})
    hp = dict(co.make_hparams(1, 64, 48, 6, 'beta-tcvae'))
    hp['beta_tcvae.beta_anneal_epochs'] = 5
    m = BetaTCVAE(hp)
    assert m.beta_vals[0] == 0 and m.kl_anneal_vals[4] == 1 and m.beta_vals[-1] == hp['beta_tcvae.beta']
    hp = dict(co.make_hparams(1, 64, 48, 6, 'vae'))
    hp['model_type'] = 'linear'
    with pytest.raises(NotImplementedError):
        VAE(hp)


@pytest.mark.parametrize('key,val', [('ae_batch_norm', True), ('ae_decoding_last_FF_layer', True)])
def test_unsupported_variants_raise_instead_of_falling_back(key, val):
    from behavenet_b200.models import AE
    hp = co.make_hparams(1, 32, 32, 8)
    hp[key] = val
    with pytest.raises(NotImplementedError):
        AE(hp)


def test_valid_padding_and_session_layers_mirror_the_reference_containers():
    """'valid' padding: zero pads, ``output_padding`` per transposed conv as aes.py:382-405 computes it, and the
    plan descriptor's negative bottom / right crop; ``fit_sess_io_layers``: module lists named like the
    reference's (aes.py:69-80, 298-312), whose entries fill the parameter table by ``dataset``."""
    from behavenet_b200.models import AE
    hp = co.make_hparams(2, 160, 130, 6, padding_type='valid', n_datasets=2)
    assert hp['ae_encoding_y_dim'] == [78, 37, 17, 7, 1] and hp['ae_encoding_x_dim'] == [63, 30, 13, 5, 1]
    m = AE(hp)
    assert set(m.state_dict()) == set(co.init_state_dict(hp))
    convs = [l for l in m.decoding.decoder if isinstance(l, torch.nn.ConvTranspose2d)]
    assert [c.output_padding for c in convs] == [(2, 0), (0, 0), (0, 1), (1, 0)]
    io, io_in = list(m.decoding.decoder)[-2], list(m.encoding.encoder)[0]
    assert isinstance(io, torch.nn.ModuleList) and len(io) == 2 and io[0].output_padding == (1, 1)
    d = m._driver.desc
    assert list(d.dec_pb)[:5] == [-2, 0, 0, -1, -1] and list(d.dec_pr)[:5] == [0, 0, -1, 0, -1]
    assert list(d.dec_pt)[:5] == [0] * 5 and list(d.enc_pb)[:5] == [0] * 5
    p0, p1 = m._kernel_params(0), m._kernel_params(1)
    assert p0[0] is io_in[0].weight and p1[0] is io_in[1].weight
    assert p0[-2] is io[0].weight and p1[-1] is io[1].bias
    assert all(a is b for a, b in zip(p0[2:-2], p1[2:-2]))
    with pytest.raises(TypeError):
        m._kernel_params(None)
    hp = co.make_hparams(1, 32, 32, 8)
    hp['ae_padding_type'] = 'circular'
    with pytest.raises(ValueError):
        AE(hp)


def test_invalid_model_type_and_linear():
    from behavenet_b200.models import AE, LinearAEDecoder
    hp = co.make_linear_hparams(2, 24, 20, 5)
    model = AE(copy.deepcopy(hp))
    # the reference's names: the decoder holds the encoder module, so its tensors are listed twice (aes.py:573)
    assert set(model.state_dict()) == set(co.init_linear_state_dict(hp))
    model.load_state_dict(co.init_linear_state_dict(hp))
    assert len(list(model.get_parameters())) == 3 and model.decoding.encoder is model.encoding
    assert 'Encoder weights transposed' in str(model)
    with pytest.raises(RuntimeError):            # no CPU path
        model(torch.rand(2, 2, 24, 20))
    with pytest.raises(RuntimeError):
        model.loss({'images': torch.rand(1, 2, 2, 24, 20)})
    copy.deepcopy(model)
    pickle.loads(pickle.dumps(model))
    with pytest.raises(NotImplementedError):
        AE(dict(hp, fit_sess_io_layers=True))
    with pytest.raises(NotImplementedError):
        AE(dict(hp, n_ae_latents=65))
    with pytest.raises(NotImplementedError):
        LinearAEDecoder(5, (2, 24, 20), None)
    hp = co.make_hparams(1, 32, 32, 8)
    hp['model_type'] = 'bogus'
    with pytest.raises((ValueError, NotImplementedError)):
        AE(hp)


def test_cpu_tensors_are_rejected():
    from behavenet_b200.models import AE
    model = AE(co.make_hparams(1, 32, 32, 8))
    with pytest.raises(RuntimeError):
        model(torch.rand(2, 1, 32, 32))
    with pytest.raises(RuntimeError):
        model.loss({'images': torch.rand(1, 2, 1, 32, 32)})


def test_hmm_m_step_matches_oracle_given_oracle_expectations():
    """Host M-step logic: feed the oracle's E-step outputs through the product's M-step (as
    Gram statistics) and compare with the oracle's M-step."""
    from behavenet_b200.ssm import HMM
    for transitions, kw in (('stationary', None), ('sticky', {'kappa': 5.0})):
        p = ao.synth_params(3, 2, 2, seed=2, mix=0.3)
        rng = np.random.RandomState(0)
        xs = [ao.sample(p, T, rng)[1] for T in (80, 60)]
        exps = ao.e_step(p, xs)
        hmm = HMM(3, 2, observations='ar', observation_kwargs={'lags': 2}, transitions=transitions,
                  transition_kwargs=kw)
        hmm.observations.As, hmm.observations.bs, hmm.observations.Sigmas = p.As.copy(), p.bs.copy(), p.Sigmas.copy()
        Sxx, Sxy, Syy, Sn = ao.ar_sufficient_stats(p, xs, exps)
        P = 2 * 2 + 1
        stats = np.zeros((3, P + 2, P + 2))
        stats[:, :P, :P], stats[:, :P, P:], stats[:, P:, P:] = Sxx, Sxy, Syy
        stats[:, P:, :P] = np.transpose(Sxy, (0, 2, 1))
        hmm.observations.m_step(stats, Sn)
        hmm.transitions.m_step(sum(e[1] for e in exps))
        hmm.init_state_distn.m_step(sum(e[0][0] for e in exps))
        ref = ao.m_step(p, xs, exps, transitions=transitions, kappa=5.0)
        used = Sn > 1          # states with < 1 expected count are re-seeded from a used one (ssm)
        assert used.sum() >= 2
        np.testing.assert_allclose(hmm.observations.As[used], ref.As[used], atol=1e-10)
        np.testing.assert_allclose(hmm.observations.bs[used], ref.bs[used], atol=1e-10)
        np.testing.assert_allclose(hmm.observations.Sigmas[used], ref.Sigmas[used], atol=1e-10)
        np.testing.assert_allclose(hmm.transitions.log_Ps, ref.log_Ps, atol=1e-10)
        np.testing.assert_allclose(hmm.init_state_distn.log_pi0, ref.log_pi0, atol=1e-10)


def test_hmm_construction_permute_pickle_and_errors():
    from behavenet_b200.ssm import HMM
    np.random.seed(1)
    hmm = HMM(4, 3, observations='ar', observation_kwargs={'lags': 2}, transitions='sticky',
              transition_kwargs={'kappa': 100})
    assert hmm.observations.As.shape == (4, 3, 6) and hmm.observations.lags == 2
    P = hmm.transitions.transition_matrix
    np.testing.assert_allclose(P.sum(1), 1, atol=1e-12)
    before = (hmm.observations.bs.copy(), hmm.transitions.log_Ps.copy())
    perm = np.array([2, 0, 3, 1])
    hmm.permute(perm)
    np.testing.assert_array_equal(hmm.observations.bs, before[0][perm])
    np.testing.assert_array_equal(hmm.transitions.log_Ps, before[1][np.ix_(perm, perm)])
    hmm.hparams = {'a': 1}
    clone = pickle.loads(pickle.dumps(hmm))
    np.testing.assert_array_equal(clone.observations.As, hmm.observations.As)
    assert clone.hparams == {'a': 1}
    for bad in (dict(observations='robust_ar'), dict(transitions='recurrent'), dict(M=2)):
        with pytest.raises(NotImplementedError):
            HMM(4, 3, **bad)
    z, x = hmm.sample(20)
    assert z.shape == (20,) and x.shape == (20, 3)


def test_shard_range_covers_everything_once():
    from behavenet_b200 import parallel
    for n in (0, 1, 7, 256, 2048, 5292):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from behavenet_b200 import parallel
    from behavenet_b200.models import AE
    assert parallel.init('gloo')
    assert parallel.world_size() == world and parallel.rank() == rank
    hp = co.make_hparams(1, 32, 32, 8)
    torch.manual_seed(0)
    model = AE(hp)
    model.data_parallel = True
    params = model._kernel_params()
    # fabricate rank-dependent gradients in the flat buffer the fused loss path would fill
    grads = model._grad_table(params)
    for i, g in enumerate(g for g in grads if g is not None):
        g.fill_(float(rank + 1) * (i + 1))
    sse = torch.tensor([1.0 + rank, 10.0], dtype=torch.float64)
    lo, hi = model._shard(10)
    model._allreduce(params, sse)
    tot = sum(r + 1 for r in range(world))
    ok = all(torch.all(g == tot * (i + 1)) for i, g in enumerate(g for g in grads if g is not None))
    ok = ok and sse.tolist() == [sum(1.0 + r for r in range(world)), 10.0 * world]
    # bucketed path: the decoder bucket's all-reduce starts early (after the decoder's backward pass), the
    # rest follows -- every gradient must still be summed exactly once
    for i, g in enumerate(g for g in grads if g is not None):
        g.fill_(float(rank + 1) * (i + 1))
    span = model._rt.bufs['flat_grad_dec']
    n_dec = sum(p.numel() for p in model.decoding.parameters())
    ok = ok and span is not None and span[1] - span[0] == n_dec and span[1] == model._rt.bufs['flat_grad'].numel()
    pending = model._allreduce_begin(params)
    ok = ok and pending is not None
    model._allreduce(params, sse.clone(), pending)
    ok = ok and all(torch.all(g == tot * (i + 1)) for i, g in enumerate(g for g in grads if g is not None))
    # PS-VAE chunks that span ranks: every rank ends up with its own rows of a per-chunk quantity that
    # depends on all rows of the chunk, and only a chunk's owner reports its sums
    from behavenet_b200.models.vaes import _latent_rows
    n_total, Lz, chunk = 23, 3, 8
    gen_ = torch.Generator().manual_seed(1)
    tab = torch.randn(n_total, 3 * Lz, generator=gen_)
    chunks = [(b, min(b + chunk, n_total)) for b in range(0, n_total, chunk)]
    beg, end = parallel.shard_range(n_total)
    owned = []

    def call(c, b, e, owner, pre_f, lv_f, eps_f, mu_f, z_f, gmu_f, glv_f, gz_f):
        sl = slice(b, e)
        mu_f[sl] = pre_f[sl] - pre_f[sl].mean(0, keepdim=True)           # needs the whole chunk
        z_f[sl] = lv_f[sl] * eps_f[sl].sum()
        if owner:
            owned.append(c)
    rows = _latent_rows(beg, end, n_total, chunks, tab[beg:end, :Lz].clone(), tab[beg:end, Lz:2 * Lz].clone(),
                        tab[beg:end, 2 * Lz:].clone(), call, True)
    want_mu = torch.cat([tab[b:e, :Lz] - tab[b:e, :Lz].mean(0, keepdim=True) for b, e in chunks])[beg:end]
    want_z = torch.cat([tab[b:e, Lz:2 * Lz] * tab[b:e, 2 * Lz:].sum() for b, e in chunks])[beg:end]
    ok = ok and torch.allclose(rows[0], want_mu, atol=1e-6) and torch.allclose(rows[1], want_z, atol=1e-5)
    cnt = torch.zeros(len(chunks))
    cnt[owned] += 1
    parallel.all_reduce_sum(cnt)
    ok = ok and cnt.tolist() == [1.0] * len(chunks)
    # second path: gradients that are NOT views of the flat buffer
    model.zero_grad()
    for p in model.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    model._allreduce(params, sse)
    ok = ok and all(torch.all(p.grad == tot) for p in model.parameters())
    # input pipeline under the same ranks: every rank stages only its own frames of each trial, all ranks
    # walk the trials in the same order, and the shards tile the trial
    import torch.distributed as dist
    from behavenet_b200.data import ArraySource, PrefetchSessionsGenerator
    rng = np.random.RandomState(0)
    src = ArraySource({'images': [rng.randint(0, 256, (T, 1, 4, 4)).astype(np.uint8) for T in (5, 8, 3, 6, 7, 4, 9, 5, 6, 7)]})
    gen = PrefetchSessionsGenerator([src], device='cpu', rng_seed=4, shard_frames=True)
    for _ in range(gen.n_tot_batches['train']):
        d, _s = gen.next_batch('train')
        idx, (beg, T) = int(d['batch_idx']), d['shard']
        n_mine = d['images'].shape[1]
        info = torch.tensor([idx, beg, n_mine, T])
        both = [torch.zeros_like(info) for _ in range(world)]
        dist.all_gather(both, info, group=parallel.group())
        ok = ok and all(int(b[0]) == idx and int(b[3]) == T == src.trial_length(idx) for b in both)
        ok = ok and int(both[0][1]) == 0 and int(both[1][1]) == int(both[0][2]) and int(both[0][2] + both[1][2]) == T
        ok = ok and np.array_equal(d['images'][0].numpy(), src.load('images', idx, beg, beg + n_mine).astype('float32') / 255)
    gen.close()
    q.put((rank, ok, (lo, hi)))
    parallel.shutdown()


def test_world_size_2_gloo_allreduce_and_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == (0, 5) and res[1][2] == (5, 10)


def test_conditional_models_protocol_cpu():
    """ConditionalAE / AEMSP: state_dict names and shapes of the reference (aes.py:776-1217), deep-copyable,
    constructor errors; no kernels are launched."""
    import copy
    from oracle import cae_oracle as co
    from behavenet_b200.models import ConditionalAE, AEMSP
    hp = co.make_hparams(2, 32, 32, 8, 'cond-ae', 4, conditional_encoder=True)
    m = ConditionalAE(copy.deepcopy(hp))
    sd = m.state_dict()
    assert sd['encoding.encoder.conv0.weight'].shape == (32, 2 + 2, 5, 5)       # frames + n_labels/2 label images
    assert sd['decoding.FF.weight'].shape[1] == 8 + 4 and sd['encoding.FF.weight'].shape[0] == 8
    assert m.hparams['hidden_layer_size'] == 12
    copy.deepcopy(m)
    hp = co.make_hparams(1, 64, 48, 6, 'cond-ae-msp', 3)
    m = AEMSP(copy.deepcopy(hp))
    assert set(m.state_dict()) == set(co.init_state_dict(hp))
    m.create_orthogonal_matrix()
    U = m.U.weight.detach().numpy()
    assert not m.U.weight.requires_grad and U.shape == (6, 6)
    P = m.projection.weight.detach().numpy()
    assert abs(U[:3] - P).max() == 0 and abs(U[3:] @ P.T).max() < 1e-6 and abs(U[3:] @ U[3:].T - np.eye(3)).max() < 1e-6
    bad = dict(hp)
    bad['n_labels'] = 7
    with pytest.raises(ValueError):
        AEMSP(bad)
    lin = dict(hp)
    lin['model_type'] = 'linear'
    with pytest.raises(NotImplementedError):
        AEMSP(lin)


def test_mspsvae_protocol_and_triplet_rule_cpu():
    """MSPSVAE (reference vaes.py:849-1462): state_dict names / frozen orthogonal A, B, C; the product's
    general triplet rule equals the oracle's term-by-term table (itself checked against the reference's
    losses.triplet_loss for 2-4 sessions by oracle/gen_golden.py); the small loss helpers equal the oracle's."""
    from behavenet_b200.models import MSPSVAE
    from behavenet_b200.models import vaes
    np.random.seed(1)
    hp = co.make_hparams(2, 32, 32, 8, 'msps-vae', 3)
    m = MSPSVAE(copy.deepcopy(hp))
    assert set(m.state_dict()) == set(co.init_state_dict(hp))
    enc = m.encoding
    assert not (enc.A.weight.requires_grad or enc.B.weight.requires_grad or enc.C.weight.requires_grad)
    assert enc.C.bias.requires_grad and enc.B.weight.shape == (8 - 3 - 2, 8)
    q = torch.cat([enc.A.weight, enc.C.weight, enc.B.weight], 0)
    assert torch.allclose(q @ q.T, torch.eye(8), atol=1e-5)
    copy.deepcopy(m)
    one = dict(hp)
    one['n_sessions_per_batch'] = 1
    with pytest.raises(ValueError):
        MSPSVAE(one)
    g = torch.Generator().manual_seed(0)
    tm = torch.nn.TripletMarginLoss(margin=1.0, p=2)
    for ns in (2, 3, 4):
        z = torch.randn(37 * ns, 4, generator=g)
        ds = np.repeat(np.arange(ns) * 3, 37)[np.random.RandomState(ns).permutation(37 * ns)]
        np.random.seed(5)
        a = vaes.triplet_loss(tm, z, ds)
        np.random.seed(5)
        b = co.triplet_loss(z, ds)
        assert abs(float(a) - float(b)) < 1e-6, ns
    with pytest.raises(NotImplementedError):
        vaes.triplet_loss(tm, torch.zeros(50, 2), np.repeat(np.arange(5), 10))
    z, mu, lv = torch.randn(9, 3, generator=g), torch.randn(9, 3, generator=g), 0.1 * torch.randn(9, 3, generator=g)
    for a, b in zip(vaes.decomposed_kl(z, mu, lv), co.decomposed_kl(z, mu, lv)):
        assert abs(float(a) - float(b)) < 1e-6
    x, xh, mk = torch.rand(5, 2, 4, 4, generator=g), torch.rand(5, 2, 4, 4, generator=g), (torch.rand(5, 2, 4, 4, generator=g) > 0.3).float()
    assert abs(float(vaes.gaussian_ll(x, xh, mk)) - float(co.gaussian_ll(x, xh, mk))) < 1e-5


def test_loss_composition_of_autograd_path_models_with_oracle_forward():
    """ConditionalAE / ConditionalVAE / AEMSP / MSPSVAE ``loss()`` bodies (chunking, weighting, dict keys) on
    CPU: the model's ``forward`` is replaced by the oracle's restatement of it, everything after that is the
    product code; values against the fixtures written by the reference classes."""
    from behavenet_b200 import models as M
    from tests.helpers import synth_cond_inputs

    def run(mc, cls, name, c, h, w, L, b, nl, chunk, fwd, n_latents=0, **loss_kw):
        gold = load_golden(name)
        hp = co.make_hparams(c, h, w, L, mc, nl)
        sd = co.init_state_dict(hp, seed=0)
        np.random.seed(0)
        model = cls(copy.deepcopy(hp))
        model.load_state_dict(sd)
        model.forward = lambda x, **kw: fwd(model, sd, hp, x, **kw)
        inp = synth_cond_inputs(c, h, w, b, nl, n_latents=n_latents)
        data = {'images': inp['x'][None], 'labels': inp['labels'][None], 'masks': inp['masks'][None]}
        if 'eps' in inp:
            loss_kw['eps'] = inp['eps']
        model.curr_epoch = 1
        out = model.loss(data, accumulate_grad=False, chunk_size=chunk, **loss_kw)
        for k, v in out.items():
            if k != 'labels_r2':
                ref = float(gold['loss.' + k])
                assert abs(v - ref) <= 1e-5 * max(1.0, abs(ref)), (name, k, v, ref)
        return out

    run('cond-ae', M.ConditionalAE, 'condae_64x48x1_l6_b7', 1, 64, 48, 6, 7, 4, 4,
        lambda m, sd, hp, x, labels=None, labels_2d=None, **kw: co.cond_ae_forward(sd, hp, x, labels, labels_2d))
    run('cond-vae', M.ConditionalVAE, 'condvae_64x48x1_l6_b7', 1, 64, 48, 6, 7, 4, 4,
        lambda m, sd, hp, x, labels=None, eps=None, **kw: co.cond_vae_forward(sd, hp, x, labels, eps), n_latents=6)
    out = run('cond-ae-msp', M.AEMSP, 'aemsp_64x48x1_l6_b7', 1, 64, 48, 6, 7, 3, 4,
              lambda m, sd, hp, x, **kw: co.aemsp_forward(sd, hp, x))
    assert set(out) == {'loss', 'loss_mse', 'loss_msp', 'labels_r2'}
    # MSPS-VAE: its fixture uses its own input draw (oracle/gen_golden.py run_reference_msps)
    gold = load_golden('mspsvae_32x32x2_l8_b24')
    hp = co.make_hparams(2, 32, 32, 8, 'msps-vae', 3)
    sd = co.init_state_dict(hp, seed=0)
    np.random.seed(0)
    model = M.MSPSVAE(copy.deepcopy(hp))
    model.load_state_dict(sd)
    model.forward = lambda x, eps=None, **kw: co.msps_forward(sd, hp, x, eps)
    g = torch.Generator().manual_seed(1234)
    x, y = torch.rand(24, 2, 32, 32, generator=g), torch.randn(24, 3, generator=g)
    eps = torch.randn(24, 8, generator=g)
    m = (torch.rand(24, 2, 32, 32, generator=g) > 0.1).float()
    model.curr_epoch = 1
    single = model.loss({'images': x[None], 'labels': y[None], 'masks': m[None]}, accumulate_grad=False, eps=eps)
    datas = [{'images': x[None, :12], 'labels': y[None, :12], 'masks': m[None, :12]},
             {'images': x[None, 12:], 'labels': y[None, 12:], 'masks': m[None, 12:]}]
    np.random.seed(7)
    multi = model.loss(datas, dataset=[0, 1], accumulate_grad=False, eps=eps)
    for tag, vals in (('single.', single), ('loss.', multi)):
        assert len(vals) == 13
        for k, v in vals.items():
            ref = float(gold[tag + k])
            assert abs(v - ref) <= 1e-5 * max(1.0, abs(ref)) + (1e-3 if k == 'label_r2' else 0), (tag, k, v, ref)
