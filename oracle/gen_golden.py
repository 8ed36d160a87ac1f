"""Generate the committed golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python oracle/gen_golden.py

* CAE / PS-VAE: imports ``behavenet.models.AE`` / ``PSVAE`` from /root/reference (with a stub for
  the absent ``commentjson`` module, which the reference only uses to read arch json files),
  loads the seeded parameters of ``oracle.cae_oracle.init_state_dict`` into them, runs
  ``forward`` and ``loss`` on seeded inputs and stores inputs' seeds + outputs.  It also asserts
  that the oracle restatement (oracle/cae_oracle.py) reproduces the reference to fp32 round-off,
  which is what pins the oracle.
* ARHMM: ``ssm`` is not available, so the fixture holds the restated oracle's own outputs on a
  small seeded problem (regression fixture, parity unpinned -- see oracle/arhmm_oracle.py).

The fixtures are small .npz files (a few hundred KB in total).
"""

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.modules.setdefault('commentjson', types.ModuleType('commentjson'))

from oracle import cae_oracle as co          # noqa: E402
from oracle import arhmm_oracle as ao        # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

# (name, C, H, W, latents, batch, model_class, n_labels, chunk_size)
CAE_CASES = [
    ('c1_ae_32x32x1_l8_b32', 1, 32, 32, 8, 32, 'ae', 0, 200),
    ('ae_64x48x1_l6_b7', 1, 64, 48, 6, 7, 'ae', 0, 4),          # integration-test geometry, 2 chunks
    ('ae_128x128x1_l12_b3', 1, 128, 128, 12, 3, 'ae', 0, 200),   # C2 geometry, tiny batch
    ('psvae_128x128x2_l16_b5', 2, 128, 128, 16, 5, 'ps-vae', 4, 3),  # C3 geometry, 2 chunks
    ('psvae_32x32x2_l8_b6', 2, 32, 32, 8, 6, 'ps-vae', 3, 200),
    ('vae_64x48x1_l6_b7', 1, 64, 48, 6, 7, 'vae', 0, 4),             # 2 chunks, beta = 2
    ('btcvae_32x32x2_l8_b6', 2, 32, 32, 8, 6, 'beta-tcvae', 0, 4),   # 2 chunks, beta = 5
    ('condae_64x48x1_l6_b7', 1, 64, 48, 6, 7, 'cond-ae', 4, 4),      # labels join the latents, 2 chunks
    ('condae_enc_32x32x2_l8_b6', 2, 32, 32, 8, 6, 'cond-ae+enc', 4, 4),   # + one-hot label images as channels
    ('aemsp_64x48x1_l6_b7', 1, 64, 48, 6, 7, 'cond-ae-msp', 3, 4),
    ('condvae_64x48x1_l6_b7', 1, 64, 48, 6, 7, 'cond-vae', 4, 4),
    ('mspsvae_32x32x2_l8_b24', 2, 32, 32, 8, 24, 'msps-vae', 3, 0),   # 2 sessions x 12 frames, triplet term
    # architecture variants of the same classes (aes.py:69-80, 298-312, 382-405): 'valid' padding (transposed convs
    # with output_padding) and per-session input / output layers (forward / loss with dataset = 1)
    ('ae_valid_128x128x1_l12_b3', 1, 128, 128, 12, 3, 'ae+valid', 0, 2),
    ('ae_valid_160x130x2_l6_b5', 2, 160, 130, 6, 5, 'ae+valid', 0, 2),
    ('ae_io3_64x48x1_l6_b7', 1, 64, 48, 6, 7, 'ae+io3', 0, 4),
    ('ae_valid_io2_160x130x2_l6_b5', 2, 160, 130, 6, 5, 'ae+valid+io2', 0, 200),
    ('psvae_valid_128x128x2_l16_b5', 2, 128, 128, 16, 5, 'ps-vae+valid', 4, 3),     # the variants reach the whole family
    ('vae_io2_64x48x1_l6_b7', 1, 64, 48, 6, 7, 'vae+io2', 0, 4),
]


def synth_inputs(case):
    name, c, h, w, L, b, mc, nl, chunk = case
    mc = mc.split('+')[0]
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(b, c, h, w, generator=g)
    out = {'x': x}
    if mc == 'ps-vae':
        out['labels'] = torch.randn(b, nl, generator=g)
        out['eps'] = torch.randn(b, L, generator=g)
    if mc in ('vae', 'beta-tcvae'):
        out['eps'] = torch.randn(b, L, generator=g)
    out['masks'] = (torch.rand(b, c, h, w, generator=g) > 0.1).float()
    return out


def run_reference_cond_vae(ConditionalVAE, vaes, name, hp, hp_ref, sd, inp, chunk):
    model = ConditionalVAE(hp_ref)
    model.load_state_dict(sd)
    model.eval()
    eps_all = inp['eps']
    state = {'pos': 0}
    orig = torch.randn_like

    def fake_randn_like(t, *a, **k):         # inject eps into reparameterize (vaes.py:33-35)
        n = t.shape[0]
        e = eps_all[state['pos']:state['pos'] + n]
        state['pos'] += n
        return e.to(t.dtype)
    vaes.torch.randn_like = fake_randn_like
    res = {}
    try:
        with torch.no_grad():
            x_hat, z, mu, logvar = model(inp['x'], labels=inp['labels'])
        res.update(x_hat=x_hat, z=z, mu=mu, logvar=logvar)
        model.curr_epoch = 1
        model.zero_grad()
        state['pos'] = 0
        loss = model.loss({'images': inp['x'][None], 'labels': inp['labels'][None], 'masks': inp['masks'][None]},
                          accumulate_grad=True, chunk_size=chunk)
    finally:
        vaes.torch.randn_like = orig
    for k, v in loss.items():
        res['loss.' + k] = torch.tensor(float(v), dtype=torch.float64)
    for k, p in model.named_parameters():
        if p.grad is not None:
            res['grad.' + k] = p.grad.clone()
    o = co.cond_vae_forward(sd, hp, inp['x'], inp['labels'], inp['eps'])
    for a, bref in zip(o, (x_hat, z, mu, logvar)):
        assert torch.allclose(a, bref, atol=2e-5), name
    lo, go = co.cond_vae_loss(sd, hp, inp['x'], inp['labels'], inp['eps'], inp['masks'], chunk_size=chunk)
    for k in lo:
        ref = float(res['loss.' + k])
        assert abs(lo[k] - ref) <= 1e-5 * max(1.0, abs(ref)), (name, k, lo[k], ref)
    assert set(go) == {k[5:] for k in res if k.startswith('grad.')}, name
    for k, gref in go.items():
        gr = res['grad.' + k]
        assert torch.allclose(gref, gr, atol=1e-4 * float(gr.abs().max()) + 1e-7), (name, k)
    return res


def run_reference_msps(case):
    """MSPSVAE (vaes.py:849-1273): forward, the single-session loss dict and the two-session loss (triplet term;
    numpy's global generator seeded before the call, as the tests do)."""
    from behavenet.models.vaes import MSPSVAE
    import behavenet.models.vaes as vaes
    import behavenet.fitting.losses as ref_losses
    name, c, h, w, L, b, mc, nl, chunk = case
    hp = co.make_hparams(c, h, w, L, mc, nl)
    sd = co.init_state_dict(hp, seed=0)
    g = torch.Generator().manual_seed(1234)
    inp = {'x': torch.rand(b, c, h, w, generator=g), 'labels': torch.randn(b, nl, generator=g),
           'eps': torch.randn(b, L, generator=g)}
    inp['masks'] = (torch.rand(b, c, h, w, generator=g) > 0.1).float()
    hp_ref = dict(hp)
    hp_ref['device'] = 'cpu'
    np.random.seed(0)
    model = MSPSVAE(hp_ref)
    model.load_state_dict(sd)
    model.eval()
    state = {'pos': 0}
    orig = torch.randn_like

    def fake_randn_like(t, *a, **k):
        n = t.shape[0]
        e = inp['eps'][state['pos']:state['pos'] + n]
        state['pos'] += n
        return e.to(t.dtype)
    vaes.torch.randn_like = fake_randn_like
    res = {}
    half = b // 2
    try:
        with torch.no_grad():
            x_hat, z, mu, logvar, y_hat = model(inp['x'])
        res.update(x_hat=x_hat, z=z, mu=mu, logvar=logvar, y_hat=y_hat)
        model.curr_epoch = 1
        state['pos'] = 0
        single = model.loss({'images': inp['x'][None], 'labels': inp['labels'][None], 'masks': inp['masks'][None]},
                            accumulate_grad=False)
        for k, v in single.items():
            res['single.' + k] = torch.tensor(float(v), dtype=torch.float64)
        datas = [{'images': inp['x'][None, :half], 'labels': inp['labels'][None, :half], 'masks': inp['masks'][None, :half]},
                 {'images': inp['x'][None, half:], 'labels': inp['labels'][None, half:], 'masks': inp['masks'][None, half:]}]
        model.zero_grad()
        state['pos'] = 0
        np.random.seed(7)
        loss = model.loss(datas, dataset=[0, 1], accumulate_grad=True)
    finally:
        vaes.torch.randn_like = orig
    for k, v in loss.items():
        res['loss.' + k] = torch.tensor(float(v), dtype=torch.float64)
    for k, p in model.named_parameters():
        if p.grad is not None:
            res['grad.' + k] = p.grad.clone()
    # pin the restatement
    o = co.msps_forward(sd, hp, inp['x'], inp['eps'])
    for a, bref in zip(o, (x_hat, z, mu, logvar, y_hat)):
        assert torch.allclose(a, bref, atol=2e-5), name
    lo, _ = co.msps_loss(sd, hp, inp['x'], inp['labels'], inp['eps'], inp['masks'], want_grads=False)
    for k in lo:
        ref = float(res['single.' + k])
        assert abs(lo[k] - ref) <= 1e-5 * max(1.0, abs(ref)), (name, 'single', k, lo[k], ref)
    sessions = np.concatenate([np.zeros(half), np.ones(b - half)])
    np.random.seed(7)
    lo, go = co.msps_loss(sd, hp, inp['x'], inp['labels'], inp['eps'], inp['masks'], sessions=sessions)
    for k in lo:
        ref = float(res['loss.' + k])
        assert abs(lo[k] - ref) <= 1e-5 * max(1.0, abs(ref)), (name, k, lo[k], ref)
    assert set(go) == {k[5:] for k in res if k.startswith('grad.')}, (name, set(go) ^ {k[5:] for k in res if k.startswith('grad.')})
    for k, gref in go.items():
        gr = res['grad.' + k]
        assert torch.allclose(gref, gr, atol=1e-4 * float(gr.abs().max()) + 1e-7), (name, k)
    # the triplet rule for 3 and 4 sessions, against the reference function itself
    tm = torch.nn.TripletMarginLoss(margin=1.0, p=2)
    for ns in (2, 3, 4):
        zz = torch.randn(40 * ns, 3, generator=g)
        ds = np.repeat(np.arange(ns), 40)[np.random.RandomState(ns).permutation(40 * ns)]
        np.random.seed(11)
        a = ref_losses.triplet_loss(tm, zz, ds)
        np.random.seed(11)
        bb = co.triplet_loss(zz, ds)
        assert abs(float(a) - float(bb)) < 1e-6, (ns, float(a), float(bb))
    return res


def run_reference_cond(case):
    """ConditionalAE / AEMSP (aes.py:776-1217) on tests.helpers.synth_cond_inputs."""
    from behavenet.models.aes import ConditionalAE, AEMSP
    from behavenet.models.vaes import ConditionalVAE
    import behavenet.models.vaes as vaes
    import importlib.util          # (a plain `import tests.helpers` would find the reference's tests package)
    spec = importlib.util.spec_from_file_location('bn_test_helpers', os.path.join(ROOT, 'tests', 'helpers.py'))
    helpers = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(helpers)
    synth_cond_inputs = helpers.synth_cond_inputs
    name, c, h, w, L, b, mc, nl, chunk = case
    cond_enc = mc.endswith('+enc')
    mc = mc.replace('+enc', '')
    hp = co.make_hparams(c, h, w, L, mc, nl, conditional_encoder=cond_enc)
    sd = co.init_state_dict(hp, seed=0)
    inp = synth_cond_inputs(c, h, w, b, nl, n_latents=L if mc == 'cond-vae' else 0)
    res = {}
    hp_ref = dict(hp)
    hp_ref['device'] = 'cpu'
    if mc == 'cond-vae':
        return run_reference_cond_vae(ConditionalVAE, vaes, name, hp, hp_ref, sd, inp, chunk)
    model = (ConditionalAE if mc == 'cond-ae' else AEMSP)(hp_ref)
    model.load_state_dict(sd)
    model.eval()
    data = {'images': inp['x'][None], 'labels': inp['labels'][None], 'masks': inp['masks'][None]}
    if cond_enc:
        data['labels_sc'] = inp['labels_2d'][None]
    with torch.no_grad():
        if mc == 'cond-ae':
            x_hat, z = model(inp['x'], labels=inp['labels'], labels_2d=inp['labels_2d'])
        else:
            x_hat, z, y = model(inp['x'])
            res['y'] = y
    res['x_hat'], res['z'] = x_hat, z
    model.zero_grad()
    loss = model.loss(data, accumulate_grad=True, chunk_size=chunk)
    for k, v in loss.items():
        res['loss.' + k] = torch.tensor(float(v), dtype=torch.float64)
    for k, p in model.named_parameters():
        if p.grad is not None:
            res['grad.' + k] = p.grad.clone()
    # pin the restatement
    if mc == 'cond-ae':
        xo, zo = co.cond_ae_forward(sd, hp, inp['x'], inp['labels'], inp['labels_2d'])
        lo, go = co.cond_ae_loss(sd, hp, inp['x'], inp['labels'], inp['labels_2d'] if cond_enc else None,
                                 inp['masks'], chunk)
    else:
        xo, zo, yo = co.aemsp_forward(sd, hp, inp['x'])
        assert torch.allclose(yo, res['y'], atol=1e-5), name
        lo, go = co.aemsp_loss(sd, hp, inp['x'], inp['labels'], inp['masks'], chunk)
    assert torch.allclose(xo, x_hat, atol=1e-6) and torch.allclose(zo, z, atol=1e-5), name
    for k in lo:
        ref = float(res['loss.' + k])
        assert abs(lo[k] - ref) <= 1e-6 * max(1.0, abs(ref)), (name, k, lo[k], ref)
    assert set(go) == {k[5:] for k in res if k.startswith('grad.')}, name
    for k, gref in go.items():
        gr = res['grad.' + k]
        assert torch.allclose(gref, gr, atol=1e-4 * float(gr.abs().max()) + 1e-8), (name, k)
    return res


def variant_options(mc):
    """'ae+valid+io3' -> ('ae', {'padding_type': 'valid', 'n_datasets': 3}): the make_hparams keywords of a case."""
    mc, *flags = mc.split('+')
    opts = {}
    for f in flags:
        if f == 'valid':
            opts['padding_type'] = 'valid'
        elif f.startswith('io'):
            opts['n_datasets'] = int(f[2:])
        else:
            raise ValueError(f)
    return mc, opts


def run_reference(case):
    from behavenet.models import AE, PSVAE
    import behavenet.models.vaes as vaes
    name, c, h, w, L, b, mc, nl, chunk = case
    if mc.startswith('cond-'):
        return run_reference_cond(case)
    if mc == 'msps-vae':
        return run_reference_msps(case)
    mc, opts = variant_options(mc)
    hp = co.make_hparams(c, h, w, L, mc, nl, **opts)
    sd = co.init_state_dict(hp, seed=0)
    inp = synth_inputs(case)
    res = {}
    hp_ref = dict(hp)
    if mc == 'ae':
        ds = 1 if opts.get('n_datasets', 0) else None       # which session's input / output layers
        model = AE(hp_ref)
        model.load_state_dict(sd)
        model.eval()
        with torch.no_grad():
            x_hat, z = model(inp['x'], dataset=ds)
        res['x_hat'], res['z'] = x_hat, z
        for tag, m in (('', None), ('_masked', inp['masks'])):
            model.zero_grad()
            data = {'images': inp['x'][None]}
            if m is not None:
                data['masks'] = m[None]
            loss = model.loss(data, dataset=ds or 0, accumulate_grad=True, chunk_size=chunk)
            res['loss' + tag] = torch.tensor(loss['loss'], dtype=torch.float64)
            for k, p in model.named_parameters():
                if p.grad is not None:                      # the other sessions' io layers stay without gradient
                    res['grad%s.%s' % (tag, k)] = p.grad.clone()
        # pin the restatement
        xo, zo = co.ae_forward(sd, hp, inp['x'], ds)
        assert torch.allclose(xo, x_hat, atol=1e-6), name
        assert torch.allclose(zo, z, atol=1e-5), name
        lo, go = co.ae_loss(sd, hp, inp['x'], None, chunk, dataset=ds)
        assert abs(lo['loss'] - float(res['loss'])) < 1e-7, name
        assert set(go) == {k[5:] for k in res if k.startswith('grad.')}, name
        for k, gref in go.items():
            assert torch.allclose(gref, res['grad.' + k], atol=1e-6, rtol=1e-4), (name, k)
    elif mc in ('vae', 'beta-tcvae'):
        from behavenet.models import VAE, BetaTCVAE
        model = (VAE if mc == 'vae' else BetaTCVAE)(hp_ref)
        model.load_state_dict(sd)
        model.eval()
        eps_all = inp['eps']
        state = {'pos': 0}
        orig = torch.randn_like

        def fake_randn_like(t, *a, **k):
            n = t.shape[0]
            e = eps_all[state['pos']:state['pos'] + n]
            state['pos'] += n
            return e.to(t.dtype)
        vaes.torch.randn_like = fake_randn_like
        try:
            ds = 1 if opts.get('n_datasets', 0) else None
            with torch.no_grad():
                state['pos'] = 0
                x_hat, z, mu, logvar = model(inp['x'], dataset=ds)
            res.update(x_hat=x_hat, z=z, mu=mu, logvar=logvar)
            model.curr_epoch = 1
            model.zero_grad()
            state['pos'] = 0
            loss = model.loss({'images': inp['x'][None]}, dataset=ds or 0, accumulate_grad=True, chunk_size=chunk)
        finally:
            vaes.torch.randn_like = orig
        for k, v in loss.items():
            res['loss.' + k] = torch.tensor(float(v), dtype=torch.float64)
        for k, p in model.named_parameters():
            if p.grad is not None:
                res['grad.' + k] = p.grad.clone()
        # pin the restatement
        o = co.vae_forward(sd, hp, inp['x'], inp['eps'], dataset=ds)
        for a, bref in zip(o, (x_hat, z, mu, logvar)):
            assert torch.allclose(a, bref, atol=2e-5), name
        if mc == 'vae':
            lo, go = co.vae_loss(sd, hp, inp['x'], inp['eps'], chunk_size=chunk, dataset=ds)
        else:
            lo, go = co.btcvae_loss(sd, hp, inp['x'], inp['eps'], chunk_size=chunk)
        assert set(go) == {k[5:] for k in res if k.startswith('grad.')}, name
        for k in lo:
            ref = float(res['loss.' + k])
            assert abs(lo[k] - ref) <= 1e-5 * max(1.0, abs(ref)), (name, k, lo[k], ref)
        for k, gref in go.items():
            gr = res['grad.' + k]
            assert torch.allclose(gref, gr, atol=1e-4 * float(gr.abs().max()) + 1e-7), (name, k)
    else:
        model = PSVAE(hp_ref)
        model.load_state_dict(sd)
        model.eval()
        eps_all = inp['eps']
        # inject eps into reparameterize (vaes.py:33-35 draws torch.randn_like)
        state = {'pos': 0}
        orig = torch.randn_like

        def fake_randn_like(t, *a, **k):
            n = t.shape[0]
            e = eps_all[state['pos']:state['pos'] + n]
            state['pos'] += n
            return e.to(t.dtype)
        vaes.torch.randn_like = fake_randn_like
        try:
            with torch.no_grad():
                state['pos'] = 0
                x_hat, z, mu, logvar, y_hat = model(inp['x'])
            res.update(x_hat=x_hat, z=z, mu=mu, logvar=logvar, y_hat=y_hat)
            model.curr_epoch = 1
            model.zero_grad()
            state['pos'] = 0
            data = {'images': inp['x'][None], 'labels': inp['labels'][None]}
            loss = model.loss(data, accumulate_grad=True, chunk_size=chunk)
        finally:
            vaes.torch.randn_like = orig
        for k, v in loss.items():
            res['loss.' + k] = torch.tensor(float(v), dtype=torch.float64)
        for k, p in model.named_parameters():
            if p.grad is not None:
                res['grad.' + k] = p.grad.clone()
        # pin the restatement
        o = co.psvae_forward(sd, hp, inp['x'], inp['eps'])
        for a, bref in zip(o, (x_hat, z, mu, logvar, y_hat)):
            assert torch.allclose(a, bref, atol=2e-5), name
        lo, go = co.psvae_loss(sd, hp, inp['x'], inp['labels'], inp['eps'], chunk_size=chunk)
        for k in lo:
            ref = float(res['loss.' + k])
            assert abs(lo[k] - ref) <= 1e-5 * max(1.0, abs(ref)), (name, k, lo[k], ref)
        for k, gref in go.items():
            gr = res['grad.' + k]
            assert torch.allclose(gref, gr, atol=1e-4 * float(gr.abs().max()) + 1e-7), (name, k)
    return res


def compact(res, big_limit=4096):
    """Store small tensors whole; for big ones store a strided sample + fp64 checksums."""
    out = {}
    for k, v in res.items():
        a = v.detach().cpu().numpy()
        if a.size <= big_limit:
            out[k] = a
        else:
            flat = a.reshape(-1)
            idx = np.linspace(0, flat.size - 1, 2048).astype(np.int64)
            out[k + '#idx'] = idx
            out[k + '#val'] = flat[idx]
            out[k + '#sum'] = np.array([flat.astype(np.float64).sum(),
                                        np.abs(flat.astype(np.float64)).sum()])
            out[k + '#shape'] = np.array(a.shape)
    return out


def gen_arhmm():
    p = ao.synth_params(K=4, D=3, lags=2, seed=3, mix=0.05)
    rng = np.random.RandomState(7)
    lengths = [40, 3, 17, 1, 64]
    xs = [ao.sample(p, T, rng)[1].astype(np.float32) for T in lengths]
    out = {'lengths': np.array(lengths), 'x': np.concatenate(xs, 0),
           'log_pi0': p.log_pi0, 'log_Ps': p.log_Ps, 'As': p.As, 'bs': p.bs, 'Sigmas': p.Sigmas,
           'lags': np.array(p.lags)}
    ez, ezz, lz, zs, lls = [], [], [], [], []
    for x in xs:
        ll = ao.ar_log_likelihoods(x, p.As, p.bs, p.Sigmas, p.lags)
        g, j, n = ao.expected_states(p.log_pi0, p.log_Ps, ll)
        ez.append(g); ezz.append(j); lz.append(n); lls.append(ll)
        zs.append(ao.viterbi(p.log_pi0, p.log_Ps, ll))
    out.update(Ez=np.concatenate(ez, 0), Ezz=np.stack(ezz), logZ=np.array(lz),
               z=np.concatenate(zs), ll=np.concatenate(lls, 0))
    np.savez_compressed(os.path.join(GOLD, 'arhmm_k4_d3_l2.npz'), **out)
    print('arhmm fixture written')


def gen_split_trials():
    """Outputs of the reference's ``split_trials`` (data/data_generator.py:42-103; the module imports the
    absent h5py at the top, which is stubbed) for the cases tests/test_data_generator.py replays."""
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))
    from behavenet.data.data_generator import split_trials
    out = {}
    cases = [(100, 0, 8, 1, 1, 0), (103, 1, 8, 1, 1, 0), (57, 5, 5, 1, 1, 1), (23, 2, 3, 2, 1, 0), (200, 7, 6, 2, 2, 2)]
    out['cases'] = np.array(cases)
    for i, (n, seed, tr, va, te, gap) in enumerate(cases):
        r = split_trials(n, rng_seed=seed, train_tr=tr, val_tr=va, test_tr=te, gap_tr=gap)
        for k in ('train', 'val', 'test'):
            out['%d_%s' % (i, k)] = np.asarray(r[k])
    np.savez_compressed(os.path.join(GOLD, 'split_trials.npz'), **out)
    print('split_trials fixture written')


LINEAR_CASES = [      # (name, C, H, W, latents, batch, chunk)
    ('linae_64x48x1_l6_b7', 1, 64, 48, 6, 7, 4),
    ('linae_40x36x2_l20_b9', 2, 40, 36, 20, 9, 200),
]


def run_reference_linear(case):
    """AE with model_type='linear' (aes.py:491-613, 684-687, 714-716): forward, loss with and without masks."""
    from behavenet.models import AE
    name, c, h, w, L, b, chunk = case
    hp = co.make_linear_hparams(c, h, w, L)
    sd = co.init_linear_state_dict(hp, seed=0)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(b, c, h, w, generator=g)
    masks = (torch.rand(b, c, h, w, generator=g) > 0.1).float()
    model = AE(dict(hp))
    model.load_state_dict(sd)
    model.eval()
    res = {}
    with torch.no_grad():
        res['x_hat'], res['z'] = model(x)
    xo, zo = co.linear_ae_forward(sd, hp, x)
    assert torch.allclose(xo, res['x_hat'], atol=1e-6) and torch.allclose(zo, res['z'], atol=1e-6), name
    for tag, m in (('', None), ('_masked', masks)):
        model.zero_grad()
        data = {'images': x[None]}
        if m is not None:
            data['masks'] = m[None]
        loss = model.loss(data, accumulate_grad=True, chunk_size=chunk)
        res['loss' + tag] = torch.tensor(loss['loss'], dtype=torch.float64)
        grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        for k, v in grads.items():
            res['grad%s.%s' % (tag, k)] = v
        lo, go = co.linear_ae_loss(sd, hp, x, m, chunk)
        assert abs(lo['loss'] - loss['loss']) < 1e-7, name
        assert set(go) == set(grads), (name, set(go) ^ set(grads))
        for k, gref in go.items():
            assert torch.allclose(gref, grads[k], atol=1e-7, rtol=1e-4), (name, k)
    return res


def main():
    os.makedirs(GOLD, exist_ok=True)
    if sys.argv[1:] == ['split_trials']:
        return gen_split_trials()
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for case in CAE_CASES:
        if only and case[0] not in only:
            continue
        res = run_reference(case)
        np.savez_compressed(os.path.join(GOLD, case[0] + '.npz'), **compact(res))
        print('wrote', case[0], {k: tuple(v.shape) for k, v in list(res.items())[:3]})
    for case in LINEAR_CASES:
        if only and case[0] not in only:
            continue
        res = run_reference_linear(case)
        np.savez_compressed(os.path.join(GOLD, case[0] + '.npz'), **compact(res))
        print('wrote', case[0])
    if not only:
        gen_arhmm()
        gen_split_trials()


if __name__ == '__main__':
    main()
