"""GPU parity of the architecture variants of the CAE path against goldens of the unmodified reference:

* ``ae_padding_type='valid'`` (reference aes.py:382-405): unpadded convolutions; the decoder's transposed
  convolutions carry ``output_padding``, which the plan sees as a negative bottom / right crop;
* ``fit_sess_io_layers`` (aes.py:69-80, 298-312): one input and one output layer per session, chosen by the
  ``dataset`` argument of ``forward`` / ``loss``.

Same tolerances as tests/test_gpu_cae.py (mode 0: fp32 round-off; mode 1: the measured TF32 bounds).
"""

import copy

import numpy as np
import pytest
import torch

from oracle import cae_oracle as co
from tests.helpers import load_golden, golden_compare, synth_inputs, rel_err
from tests.test_gpu_cae import tols, compare_grad

pytestmark = pytest.mark.gpu

VARIANTS = {
    # name: (C, H, W, latents, batch, make_hparams keywords, chunk)
    'ae_valid_128x128x1_l12_b3': (1, 128, 128, 12, 3, dict(padding_type='valid'), 2),
    'ae_valid_160x130x2_l6_b5': (2, 160, 130, 6, 5, dict(padding_type='valid'), 2),
    'ae_io3_64x48x1_l6_b7': (1, 64, 48, 6, 7, dict(n_datasets=3), 4),
    'ae_valid_io2_160x130x2_l6_b5': (2, 160, 130, 6, 5, dict(padding_type='valid', n_datasets=2), 200),
}


def build(case, tc_mode):
    from behavenet_b200 import _lib
    from behavenet_b200.models import AE
    c, h, w, L, b, kw, chunk = VARIANTS[case]
    hp = co.make_hparams(c, h, w, L, 'ae', 0, **kw)
    sd = co.init_state_dict(hp, seed=0)
    model = AE(copy.deepcopy(hp))
    model.load_state_dict(sd)
    model.to('cuda')
    _lib.lib().bn_set_tensor_core_mode(tc_mode)
    return model, hp, sd, synth_inputs(c, h, w, L, b, 0), chunk, (1 if kw.get('n_datasets') else None)


@pytest.mark.parametrize('tc_mode', [0, 1])
@pytest.mark.parametrize('case', list(VARIANTS))
def test_variant_forward_and_loss_match_reference_goldens(case, tc_mode):
    model, hp, sd, inp, chunk, ds = build(case, tc_mode)
    gold = load_golden(case)
    t = tols(tc_mode)
    x = inp['x'].cuda()
    with torch.no_grad():
        x_hat, z = model(x, dataset=ds)
    assert x_hat.shape == x.shape and z.shape == (x.shape[0], hp['n_ae_latents'])
    golden_compare(gold, 'x_hat', x_hat, rtol=t['xhat'], atol=t['xhat'])
    assert rel_err(z, gold['z']) < t['z'] * 10
    for tag, m in (('', None), ('_masked', inp['masks'].cuda())):
        model.zero_grad()
        data = {'images': x[None]}
        if m is not None:
            data['masks'] = m[None]
        out = model.loss(data, dataset=ds or 0, accumulate_grad=True, chunk_size=chunk)
        ref = float(gold['loss' + tag])
        assert abs(out['loss'] - ref) <= t['loss'] * abs(ref), (out, ref)
        n_checked, bad = 0, []
        for name, p in model.named_parameters():
            key = 'grad%s.%s' % (tag, name)
            if key in gold or key + '#val' in gold:
                # mode 0, 'valid' geometry: the 1 x 1 bottleneck leaves the encoder-side gradients of these 3-5 frame
                # batches as 1e-9-sized sums of cancelling terms (the output side is 1e-2); fp32 summation order
                # alone moves them by ~2e-3 of their largest entry (measured), hence the factor there
                loose = tc_mode == 0 and 'valid' in case and name.startswith('encoding.encoder')
                try:
                    compare_grad(gold, key, p.grad, tc_mode, t, factor=50.0 if loose else 1.0)
                except AssertionError as e:
                    bad.append((key, str(e)[:300]))
                n_checked += 1
            else:       # the other sessions' input / output layers took no part in the pass
                assert '_sess_io_layers.' in name and p.grad is None, name
        assert not bad, bad
        assert n_checked == 4 * len(hp['ae_encoding_n_channels']) + 4


def test_session_layers_follow_the_dataset_argument():
    """Every session's layers against the CPU oracle (mode 0), the autograd bridge included; a missing dataset
    index fails like the reference's ``layer[dataset]``."""
    case = 'ae_io3_64x48x1_l6_b7'
    model, hp, sd, inp, chunk, _ = build(case, 0)
    x = inp['x'].cuda()
    outs = []
    for ds in range(3):
        with torch.no_grad():
            x_hat, z = model(x, dataset=ds)
        xo, zo = co.ae_forward(sd, hp, inp['x'], ds)
        assert float((x_hat.cpu() - xo).abs().max()) < 1e-4 and rel_err(z, zo.numpy()) < 2e-4
        outs.append(x_hat)
        model.zero_grad()
        out = model.loss({'images': x[None]}, dataset=ds, accumulate_grad=True, chunk_size=chunk)
        lo, go = co.ae_loss(sd, hp, inp['x'], None, chunk, dataset=ds)
        assert abs(out['loss'] - lo['loss']) <= 1e-5 * lo['loss']
        got = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        assert set(got) == set(go)
        for k, g in go.items():
            scale = float(g.abs().max())
            assert float((got[k].cpu() - g).abs().max()) <= 1e-4 * scale + 1e-9, (ds, k)
        # the differentiable forward (EncodeFn / DecodeFn) takes the same session's layers
        model.zero_grad()
        x_hat, _ = model(x[:3], dataset=ds)
        ((x_hat - x[:3]) ** 2).mean().backward()
        _, go3 = co.ae_loss(sd, hp, inp['x'][:3], None, 200, dataset=ds)
        got = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        assert set(got) == set(go3)
        for k, g in go3.items():
            assert float((got[k].cpu() - g).abs().max()) <= 1e-4 * float(g.abs().max()) + 1e-9, (ds, k)
    assert float((outs[0] - outs[1]).abs().max()) > 1e-3 and float((outs[1] - outs[2]).abs().max()) > 1e-3
    with pytest.raises(TypeError):
        model(x, dataset=None)


@pytest.mark.parametrize('mode', [0, 1])
def test_valid_padding_at_a_tile_filling_batch(mode):
    """The 'valid' geometry at a batch that fills the tensor-core tiles (two chunks, 32 + 8 frames) against the
    fp64 oracle, with per-parameter bounds derived from the full-size tests (tests/test_gpu_cae_fullsize.py)."""
    from behavenet_b200 import _lib
    from behavenet_b200.models import AE
    from tests.test_gpu_cae_fullsize import BOUND_FP32, BOUND_TF32, bound, f64
    hp = co.make_hparams(1, 128, 128, 12, 'ae', 0, padding_type='valid')
    sd = co.init_state_dict(hp, seed=1)
    x = torch.rand(40, 1, 128, 128, generator=torch.Generator().manual_seed(5))
    l64, g64 = co.ae_loss(f64(sd), hp, x.double(), None, chunk_size=32)
    model = AE(copy.deepcopy(hp))
    model.load_state_dict(sd)
    model.cuda()
    _lib.lib().bn_set_tensor_core_mode(mode)
    try:
        out = model.loss({'images': x.cuda()[None]}, accumulate_grad=True, chunk_size=32)
    finally:
        _lib.lib().bn_set_tensor_core_mode(1)
    assert abs(out['loss'] - l64['loss']) <= (2e-6 if mode == 0 else 1e-5) * l64['loss']
    # twice the B = 256 bounds: 40 frames average the operand-rounding noise over six times fewer terms
    # (measured here: 1.05x .. 1.6x the B = 256 bounds on the decoder's weight gradients)
    table = BOUND_FP32 if mode == 0 else BOUND_TF32
    errs = {k: rel_err(p.grad, g64[k]) for k, p in model.named_parameters()}
    bad = {k: e for k, e in errs.items() if not e < 2 * bound(table, k)}
    assert not bad, (bad, errs)


def _family_grads(gold, model, tc_mode, t, loose_prefix=None):
    bad, n_checked = [], 0
    for name, p in model.named_parameters():
        key = 'grad.' + name
        if key not in gold and key + '#val' not in gold:
            # frozen PS-VAE projections, or another session's input / output layers
            assert (not p.requires_grad) or ('_sess_io_layers.' in name and p.grad is None), name
            continue
        loose = loose_prefix is not None and tc_mode == 0 and name.startswith(loose_prefix)
        try:
            compare_grad(gold, key, p.grad, tc_mode, t, factor=50.0 if loose else 4.0)
        except AssertionError as e:
            bad.append((key, str(e)[:300]))
        n_checked += 1
    assert not bad, bad
    return n_checked


@pytest.mark.parametrize('tc_mode', [0, 1])
def test_psvae_with_valid_padding_matches_reference_golden(tc_mode):
    """The variants reach the whole model family: PS-VAE (C3 geometry) with ae_padding_type='valid'."""
    from behavenet_b200 import _lib
    from behavenet_b200.models import PSVAE
    case = 'psvae_valid_128x128x2_l16_b5'
    hp = co.make_hparams(2, 128, 128, 16, 'ps-vae', 4, padding_type='valid')
    model = PSVAE(copy.deepcopy(hp))
    model.load_state_dict(co.init_state_dict(hp, seed=0))
    model.to('cuda')
    _lib.lib().bn_set_tensor_core_mode(tc_mode)
    try:
        inp = synth_inputs(2, 128, 128, 16, 5, 4)
        gold, t = load_golden(case), tols(tc_mode)
        x, y, eps = inp['x'].cuda(), inp['labels'].cuda(), inp['eps'].cuda()
        with torch.no_grad():
            x_hat, z, mu, logvar, y_hat = model(x, eps=eps)
        golden_compare(gold, 'x_hat', x_hat, rtol=t['xhat'], atol=t['xhat'])
        for key, val in (('z', z), ('mu', mu), ('logvar', logvar), ('y_hat', y_hat)):
            assert rel_err(val, gold[key]) < t['z'] * 10, key
        model.curr_epoch = 1
        model.zero_grad()
        out = model.loss({'images': x[None], 'labels': y[None]}, accumulate_grad=True, chunk_size=3, eps=eps)
        for k in ['loss', 'loss_data_ll', 'loss_label_ll', 'loss_zs_kl', 'loss_zu_mi', 'loss_zu_tc', 'loss_zu_dwkl',
                  'loss_data_mse']:
            ref = float(gold['loss.' + k])
            assert abs(out[k] - ref) <= 10 * t['loss'] * max(1.0, abs(ref)), (k, out[k], ref)
        assert _family_grads(gold, model, tc_mode, t, loose_prefix='encoding.encoder') >= 26
    finally:
        _lib.lib().bn_set_tensor_core_mode(1)


@pytest.mark.parametrize('tc_mode', [0, 1])
def test_vae_with_session_layers_matches_reference_golden(tc_mode):
    """VAE with fit_sess_io_layers: forward / loss with dataset = 1 take session 1's input and output layers."""
    from behavenet_b200 import _lib
    from behavenet_b200.models import VAE
    case = 'vae_io2_64x48x1_l6_b7'
    hp = co.make_hparams(1, 64, 48, 6, 'vae', 0, n_datasets=2)
    model = VAE(copy.deepcopy(hp))
    model.load_state_dict(co.init_state_dict(hp, seed=0))
    model.to('cuda')
    _lib.lib().bn_set_tensor_core_mode(tc_mode)
    try:
        inp = synth_inputs(1, 64, 48, 6, 7, variational=True)
        gold, t = load_golden(case), tols(tc_mode)
        x, eps = inp['x'].cuda(), inp['eps'].cuda()
        with torch.no_grad():
            x_hat, z, mu, logvar = model(x, dataset=1, eps=eps)
        golden_compare(gold, 'x_hat', x_hat, rtol=t['xhat'], atol=t['xhat'])
        for key, val in (('z', z), ('mu', mu), ('logvar', logvar)):
            assert rel_err(val, gold[key]) < t['z'] * 10, key
        model.curr_epoch = 1
        model.zero_grad()
        out = model.loss({'images': x[None]}, dataset=1, accumulate_grad=True, chunk_size=4, eps=eps)
        for k in [k[5:] for k in gold if k.startswith('loss.')]:
            ref = float(gold['loss.' + k])
            assert abs(out[k] - ref) <= 10 * t['loss'] * max(1.0, abs(ref)), (k, out[k], ref)
        assert _family_grads(gold, model, tc_mode, t) == 26
    finally:
        _lib.lib().bn_set_tensor_core_mode(1)
