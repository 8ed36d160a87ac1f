"""CPU tests that pin the CAE oracle: against the committed golden vectors (generated from the
UNMODIFIED reference classes by oracle/gen_golden.py), against the closed forms the reference's own
tests use (tests/test_fitting/test_losses.py), and the layer-geometry table of
tests/test_models/test_ae_model_architecture_generator.py:374-405."""

import math

import numpy as np
import pytest
import torch

from oracle import cae_oracle as co
from behavenet_b200.models import ae_model_architecture_generator as gen
from tests.helpers import load_golden, golden_compare, synth_inputs

LN2PI = math.log(2 * math.pi)

CASES = {
    'c1_ae_32x32x1_l8_b32': (1, 32, 32, 8, 32, 'ae', 0, 200),
    'ae_64x48x1_l6_b7': (1, 64, 48, 6, 7, 'ae', 0, 4),
    'ae_128x128x1_l12_b3': (1, 128, 128, 12, 3, 'ae', 0, 200),
    'psvae_128x128x2_l16_b5': (2, 128, 128, 16, 5, 'ps-vae', 4, 3),
    'psvae_32x32x2_l8_b6': (2, 32, 32, 8, 6, 'ps-vae', 3, 200),
}


@pytest.mark.parametrize('case', list(CASES))
def test_oracle_reproduces_reference_goldens(case):
    c, h, w, L, b, mc, nl, chunk = CASES[case]
    gold = load_golden(case)
    hp = co.make_hparams(c, h, w, L, mc, nl)
    sd = co.init_state_dict(hp, seed=0)
    inp = synth_inputs(c, h, w, L, b, nl)
    torch.set_num_threads(8)
    if mc == 'ae':
        x_hat, z = co.ae_forward(sd, hp, inp['x'])
        golden_compare(gold, 'x_hat', x_hat, rtol=1e-5, atol=1e-6)
        golden_compare(gold, 'z', z, rtol=1e-4, atol=1e-5)
        for tag, m in (('', None), ('_masked', inp['masks'])):
            loss, grads = co.ae_loss(sd, hp, inp['x'], m, chunk)
            assert abs(loss['loss'] - float(gold['loss' + tag])) < 1e-7
            for k, g in grads.items():
                key = 'grad%s.%s' % (tag, k)
                scale = float(np.abs(gold[key + '#val'] if key + '#val' in gold else gold[key]).max())
                golden_compare(gold, key, g, rtol=1e-4, atol=1e-5 * scale + 1e-9)
    else:
        out = co.psvae_forward(sd, hp, inp['x'], inp['eps'])
        for key, val in zip(('x_hat', 'z', 'mu', 'logvar', 'y_hat'), out):
            golden_compare(gold, key, val, rtol=1e-4, atol=2e-5)
        loss, grads = co.psvae_loss(sd, hp, inp['x'], inp['labels'], inp['eps'], chunk_size=chunk)
        for k, v in loss.items():
            ref = float(gold['loss.' + k])
            assert abs(v - ref) <= 1e-5 * max(1.0, abs(ref)), k
        for k, g in grads.items():
            key = 'grad.' + k
            scale = float(np.abs(gold[key + '#val'] if key + '#val' in gold else gold[key]).max())
            golden_compare(gold, key, g, rtol=1e-3, atol=1e-4 * scale + 1e-9)


VARIANT_CASES = {     # name: (C, H, W, latents, batch, make_hparams keywords, chunk, dataset)
    'ae_valid_128x128x1_l12_b3': (1, 128, 128, 12, 3, dict(padding_type='valid'), 2, None),
    'ae_valid_160x130x2_l6_b5': (2, 160, 130, 6, 5, dict(padding_type='valid'), 2, None),
    'ae_io3_64x48x1_l6_b7': (1, 64, 48, 6, 7, dict(n_datasets=3), 4, 1),
    'ae_valid_io2_160x130x2_l6_b5': (2, 160, 130, 6, 5, dict(padding_type='valid', n_datasets=2), 200, 1),
}


@pytest.mark.parametrize('case', list(VARIANT_CASES))
def test_oracle_reproduces_reference_goldens_of_arch_variants(case):
    """'valid' padding (aes.py:382-405) and per-session input / output layers (aes.py:69-80, 298-312)."""
    c, h, w, L, b, kw, chunk, ds = VARIANT_CASES[case]
    gold = load_golden(case)
    hp = co.make_hparams(c, h, w, L, 'ae', 0, **kw)
    sd = co.init_state_dict(hp, seed=0)
    inp = synth_inputs(c, h, w, L, b, 0)
    torch.set_num_threads(8)
    x_hat, z = co.ae_forward(sd, hp, inp['x'], ds)
    golden_compare(gold, 'x_hat', x_hat, rtol=1e-5, atol=1e-6)
    golden_compare(gold, 'z', z, rtol=1e-4, atol=1e-5)
    for tag, m in (('', None), ('_masked', inp['masks'])):
        loss, grads = co.ae_loss(sd, hp, inp['x'], m, chunk, dataset=ds)
        assert abs(loss['loss'] - float(gold['loss' + tag])) < 1e-7
        stored = {k.split('#')[0] for k in gold if k.startswith('grad%s.' % tag)}
        assert stored == {'grad%s.%s' % (tag, k) for k in grads}      # only the chosen session's io layers
        for k, g in grads.items():
            key = 'grad%s.%s' % (tag, k)
            scale = float(np.abs(gold[key + '#val'] if key + '#val' in gold else gold[key]).max())
            golden_compare(gold, key, g, rtol=1e-4, atol=1e-5 * scale + 1e-9)


def test_oracle_reproduces_family_variant_goldens():
    """PS-VAE with 'valid' padding and VAE with per-session layers (dataset = 1): loss dicts of the reference."""
    torch.set_num_threads(8)
    gold = load_golden('psvae_valid_128x128x2_l16_b5')
    hp = co.make_hparams(2, 128, 128, 16, 'ps-vae', 4, padding_type='valid')
    sd = co.init_state_dict(hp, seed=0)
    inp = synth_inputs(2, 128, 128, 16, 5, 4)
    out = co.psvae_forward(sd, hp, inp['x'], inp['eps'])
    for key, val in zip(('x_hat', 'z', 'mu', 'logvar', 'y_hat'), out):
        golden_compare(gold, key, val, rtol=1e-4, atol=2e-5)
    loss, grads = co.psvae_loss(sd, hp, inp['x'], inp['labels'], inp['eps'], chunk_size=3)
    for k, v in loss.items():
        ref = float(gold['loss.' + k])
        assert abs(v - ref) <= 1e-5 * max(1.0, abs(ref)), k
    gold = load_golden('vae_io2_64x48x1_l6_b7')
    hp = co.make_hparams(1, 64, 48, 6, 'vae', 0, n_datasets=2)
    sd = co.init_state_dict(hp, seed=0)
    inp = synth_inputs(1, 64, 48, 6, 7, variational=True)
    loss, grads = co.vae_loss(sd, hp, inp['x'], inp['eps'], chunk_size=4, dataset=1)
    for k in [k[5:] for k in gold if k.startswith('loss.')]:
        ref = float(gold['loss.' + k])
        assert abs(loss[k] - ref) <= 1e-5 * max(1.0, abs(ref)), k
    assert {'grad.' + k for k in grads} == {k.split('#')[0] for k in gold if k.startswith('grad.')}
    assert any('_sess_io_layers.1.' in k for k in grads) and not any('_sess_io_layers.0.' in k for k in grads)


LINEAR_CASES = {'linae_64x48x1_l6_b7': (1, 64, 48, 6, 7, 4), 'linae_40x36x2_l20_b9': (2, 40, 36, 20, 9, 200)}


@pytest.mark.parametrize('case', list(LINEAR_CASES))
def test_oracle_reproduces_reference_goldens_of_the_linear_ae(case):
    """model_type='linear' (aes.py:491-613): decoder on the transposed encoder weights + its own bias."""
    c, h, w, L, b, chunk = LINEAR_CASES[case]
    gold = load_golden(case)
    hp = co.make_linear_hparams(c, h, w, L)
    sd = co.init_linear_state_dict(hp, seed=0)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(b, c, h, w, generator=g)
    masks = (torch.rand(b, c, h, w, generator=g) > 0.1).float()
    x_hat, z = co.linear_ae_forward(sd, hp, x)
    golden_compare(gold, 'x_hat', x_hat, rtol=1e-5, atol=1e-6)
    golden_compare(gold, 'z', z, rtol=1e-5, atol=1e-6)
    for tag, m in (('', None), ('_masked', masks)):
        loss, grads = co.linear_ae_loss(sd, hp, x, m, chunk)
        assert abs(loss['loss'] - float(gold['loss' + tag])) < 1e-7
        assert {'grad%s.%s' % (tag, k) for k in grads} == {k.split('#')[0] for k in gold if k.startswith('grad%s.' % tag)}
        for k, gr in grads.items():
            key = 'grad%s.%s' % (tag, k)
            scale = float(np.abs(gold[key + '#val'] if key + '#val' in gold else gold[key]).max())
            golden_compare(gold, key, gr, rtol=1e-4, atol=1e-6 * scale + 1e-10)


VAE_CASES = {
    'vae_64x48x1_l6_b7': (1, 64, 48, 6, 7, 'vae', 4),
    'btcvae_32x32x2_l8_b6': (2, 32, 32, 8, 6, 'beta-tcvae', 4),
}


@pytest.mark.parametrize('case', list(VAE_CASES))
def test_oracle_reproduces_reference_vae_goldens(case):
    """VAE (vaes.py:38-208) and beta-TC-VAE (vaes.py:367-503) goldens written by the reference classes."""
    c, h, w, L, b, mc, chunk = VAE_CASES[case]
    gold = load_golden(case)
    hp = co.make_hparams(c, h, w, L, mc)
    sd = co.init_state_dict(hp, seed=0)
    inp = synth_inputs(c, h, w, L, b, variational=True)
    torch.set_num_threads(8)
    out = co.vae_forward(sd, hp, inp['x'], inp['eps'])
    for key, val in zip(('x_hat', 'z', 'mu', 'logvar'), out):
        golden_compare(gold, key, val, rtol=1e-4, atol=2e-5)
    fn = co.vae_loss if mc == 'vae' else co.btcvae_loss
    loss, grads = fn(sd, hp, inp['x'], inp['eps'], chunk_size=chunk)
    for k, v in loss.items():
        ref = float(gold['loss.' + k])
        assert abs(v - ref) <= 1e-5 * max(1.0, abs(ref)), k
    for k, g in grads.items():
        key = 'grad.' + k
        scale = float(np.abs(gold[key + '#val'] if key + '#val' in gold else gold[key]).max())
        golden_compare(gold, key, g, rtol=1e-3, atol=1e-4 * scale + 1e-9)


def test_mse_closed_forms():
    x = torch.rand(5, 3)
    assert co.mse(x, x) == 0
    a = torch.tensor([1., 2, 3, 4, 5, 6])
    b = torch.tensor([2., 3, 4, 5, 6, 7])
    m = torch.tensor([1., 0, 1, 0, 1, 0])
    assert co.mse(a, b, m) == 0.5


def test_gaussian_ll_closed_forms():
    n_batch, n_dims = 5, 3
    x = torch.rand(n_batch, n_dims)
    assert abs(float(co.gaussian_ll(x, x)) + 0.5 * LN2PI * n_dims) < 1e-6
    m = torch.zeros(n_batch, n_dims)
    m[:, 0] = 1
    ll = co.gaussian_ll(torch.ones(n_batch, n_dims), torch.zeros(n_batch, n_dims), masks=m)
    assert abs(float(ll) - (-0.5 * LN2PI * n_dims - 0.5)) < 1e-6
    ll = co.gaussian_ll(torch.ones(n_batch, n_dims), torch.zeros(n_batch, n_dims))
    mse_expected = 2 * (-float(ll) - 0.5 * LN2PI * n_dims) / n_dims
    assert abs(co.gaussian_ll_to_mse(float(ll), n_dims) - mse_expected) < 1e-6


def test_kl_and_decomposed_kl():
    assert float(co.kl_div_to_std_normal(torch.zeros(1, 1), torch.zeros(1, 1))) == 0
    g = torch.Generator().manual_seed(0)
    z, mu, lv = (torch.randn(7, 4, generator=g) for _ in range(3))
    mi, tc, dw = co.decomposed_kl(z, mu, lv)
    # the three terms telescope to mean_j [log q(z_j | x_j) - log p(z_j)]
    lq = (-0.5 * (torch.exp(-lv) * (z - mu) ** 2 + lv + LN2PI)).sum(1)
    lp = (-0.5 * (z ** 2 + LN2PI)).sum(1)
    assert abs(float(mi + tc + dw) - float((lq - lp).mean())) < 1e-5


def test_handcrafted_dims_match_reference_table():
    arch = gen.load_default_arch()
    arch['ae_input_dim'] = [2, 128, 128]
    arch = gen.get_handcrafted_dims(arch, symmetric=True)
    assert arch['ae_encoding_x_dim'] == [64, 32, 16, 8, 2]
    assert arch['ae_encoding_y_dim'] == [64, 32, 16, 8, 2]
    assert arch['ae_encoding_x_padding'] == [(1, 2), (1, 2), (1, 2), (1, 2), (1, 1)]
    assert arch['ae_encoding_y_padding'] == [(1, 2), (1, 2), (1, 2), (1, 2), (1, 1)]
    assert arch['ae_decoding_x_dim'] == [8, 16, 32, 64, 128]
    assert arch['ae_decoding_y_dim'] == [8, 16, 32, 64, 128]
    assert arch['ae_decoding_x_padding'] == [(1, 1), (1, 2), (1, 2), (1, 2), (1, 2)]
    assert arch['ae_decoding_n_channels'] == [256, 128, 64, 32, 2]
    assert arch['ae_decoding_starting_dim'] == [512, 2, 2]
    # asymmetric decoder, as in the reference test
    arch1 = gen.load_default_arch()
    arch1['ae_input_dim'] = [2, 128, 128]
    arch1['ae_decoding_n_channels'] = [64, 32, 32]
    arch1['ae_decoding_kernel_size'] = [5, 5, 5]
    arch1['ae_decoding_stride_size'] = [2, 2, 2]
    arch1['ae_decoding_layer_type'] = ['conv', 'conv', 'conv']
    arch1['ae_decoding_starting_dim'] = [1, 8, 8]
    arch1 = gen.get_handcrafted_dims(arch1, symmetric=False)
    assert arch1['ae_decoding_x_dim'] == [15, 29, 57]
    assert arch1['ae_decoding_x_padding'] == [(2, 2), (2, 2), (2, 2)]
    # integration-test geometry (64x48): asymmetric last-layer pads
    arch2 = gen.load_default_arch()
    arch2['ae_input_dim'] = [1, 64, 48]
    arch2 = gen.get_handcrafted_dims(arch2)
    assert arch2['ae_encoding_y_dim'] == [32, 16, 8, 4, 1]
    assert arch2['ae_encoding_x_dim'] == [24, 12, 6, 3, 1]
    assert arch2['ae_encoding_y_padding'][-1] == (0, 1) and arch2['ae_encoding_x_padding'][-1] == (1, 1)


@pytest.mark.parametrize('inp,k,s,expect', [
    (16, 4, 3, (6, 1, 2)), (17, 4, 3, (6, 1, 1)), (16, 3, 2, (8, 0, 1)), (17, 3, 2, (9, 1, 1)),
    (16, 5, 2, (8, 1, 2)), (8, 5, 5, (2, 1, 1)), (2, 5, 5, (1, 1, 2))])
def test_calculate_output_dim_same(inp, k, s, expect):
    assert gen.calculate_output_dim(inp, k, s, 'same', 'conv') == expect


def test_calculate_output_dim_valid_and_errors():
    assert gen.calculate_output_dim(16, 4, 3, 'valid', 'conv') == (5, 0, 0)
    assert gen.calculate_output_dim(16, 2, 2, 'same', 'maxpool') == (8, 0, 0)
    assert gen.calculate_output_dim(17, 2, 2, 'valid', 'maxpool') == (8, 0, 0)
    with pytest.raises(NotImplementedError):
        gen.calculate_output_dim(16, 3, 2, 'same', 'maxpool')
    with pytest.raises(NotImplementedError):
        gen.calculate_output_dim(16, 3, 2, 'test', 'conv')
    with pytest.raises(NotImplementedError):
        gen.calculate_output_dim(16, 3, 2, 'same', 'test')
