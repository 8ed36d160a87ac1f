"""GPU parity of the linear autoencoder (reference aes.py:491-613: ``model_type='linear'``, decoder tied to the
encoder's transposed weights) against goldens of the unmodified reference and against the CPU oracle at a
full-size batch.  fp32 CUDA-core kernels: round-off tolerances only."""

import copy

import numpy as np
import pytest
import torch

from oracle import cae_oracle as co
from tests.helpers import load_golden, golden_compare, rel_err

pytestmark = pytest.mark.gpu

CASES = {   # name: (C, H, W, latents, batch, chunk)
    'linae_64x48x1_l6_b7': (1, 64, 48, 6, 7, 4),
    'linae_40x36x2_l20_b9': (2, 40, 36, 20, 9, 200),
}


def _inputs(c, h, w, b):
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(b, c, h, w, generator=g)
    return x, (torch.rand(b, c, h, w, generator=g) > 0.1).float()


def _model(hp, sd):
    from behavenet_b200.models import AE
    model = AE(copy.deepcopy(hp))
    model.load_state_dict(sd)
    return model.cuda()


@pytest.mark.parametrize('case', list(CASES))
def test_linear_ae_matches_reference_goldens(case):
    c, h, w, L, b, chunk = CASES[case]
    hp = co.make_linear_hparams(c, h, w, L)
    sd = co.init_linear_state_dict(hp, seed=0)
    gold = load_golden(case)
    model = _model(hp, sd)
    x, masks = _inputs(c, h, w, b)
    x, masks = x.cuda(), masks.cuda()
    with torch.no_grad():
        x_hat, z = model(x)
    assert x_hat.shape == x.shape and z.shape == (b, L)
    golden_compare(gold, 'x_hat', x_hat, rtol=1e-5, atol=1e-5)
    golden_compare(gold, 'z', z, rtol=1e-5, atol=1e-5)
    zz, none1, none2 = model.encoding(x)
    assert none1 is None and none2 is None and torch.equal(zz, z)
    assert torch.equal(model.decoding(z), x_hat)
    for tag, m in (('', None), ('_masked', masks)):
        model.zero_grad()
        data = {'images': x[None]}
        if m is not None:
            data['masks'] = m[None]
        out = model.loss(data, accumulate_grad=True, chunk_size=chunk)
        ref = float(gold['loss' + tag])
        assert abs(out['loss'] - ref) <= 1e-5 * ref, (out, ref)
        for name, p in model.named_parameters():
            key = 'grad%s.%s' % (tag, name)
            scale = float(np.abs(gold[key + '#val'] if key + '#val' in gold else gold[key]).max())
            golden_compare(gold, key, p.grad, rtol=1e-4, atol=1e-4 * scale)
        # loss only: same value, gradients untouched
        before = {k: p.grad.clone() for k, p in model.named_parameters()}
        again = model.loss(data, accumulate_grad=False, chunk_size=chunk)
        assert abs(again['loss'] - out['loss']) <= 1e-6 * ref
        assert all(torch.equal(before[k], p.grad) for k, p in model.named_parameters())


@pytest.mark.parametrize('L', [12, 40])
def test_linear_ae_full_size_batch_and_frame_shards(L):
    """300 frames of 128 x 128 (two reference chunks, 200 + 100) against the fp64 oracle; the same batch as three
    frame shards with data['shard'] accumulates to the same gradients (what a data-parallel all-reduce sums)."""
    from behavenet_b200 import parallel
    hp = co.make_linear_hparams(1, 128, 128, L)
    sd = co.init_linear_state_dict(hp, seed=2)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(300, 1, 128, 128, generator=g)
    m = (torch.rand(300, 1, 128, 128, generator=g) > 0.2).float()
    lo, go = co.linear_ae_loss({k: v.double() for k, v in sd.items()}, hp, x.double(), m.double(), 200)
    model = _model(hp, sd)
    xg, mg = x.cuda(), m.cuda()
    out = model.loss({'images': xg[None], 'masks': mg[None]})
    assert abs(out['loss'] - lo['loss']) <= 2e-6 * lo['loss']
    full = {k: p.grad.clone() for k, p in model.named_parameters()}
    for k, gref in go.items():
        assert rel_err(full[k], gref) < 2e-4, k
    model.zero_grad()
    total = 0.0
    for r in range(3):
        b, e = parallel.shard_range(300, 3, r)
        total += model.loss({'images': xg[b:e][None], 'masks': mg[b:e][None], 'shard': (b, 300)})['loss']
    assert abs(total - out['loss']) <= 1e-6 * out['loss']
    for k, p in model.named_parameters():
        assert rel_err(p.grad, full[k].cpu().numpy()) < 2e-4, k


def test_linear_ae_adam_trajectory_tracks_the_oracle():
    """15 Adam(amsgrad) steps (the optimizer of fitting/training.py:284-286) on structured frames against the same
    steps taken by the CPU oracle: the loss curves stay together and go down."""
    hp = co.make_linear_hparams(1, 32, 32, 8)
    sd = co.init_linear_state_dict(hp, seed=1)
    g = torch.Generator().manual_seed(3)
    basis = torch.rand(4, 1, 32, 32, generator=g)
    x = (torch.rand(64, 4, generator=g) @ basis.view(4, -1)).view(64, 1, 32, 32) / 4 + 0.02 * torch.rand(64, 1, 32, 32, generator=g)
    model = _model(hp, sd)
    opt = torch.optim.Adam(model.get_parameters(), lr=2e-3, amsgrad=True)
    names = ['encoding.encoder.weight', 'encoding.encoder.bias', 'decoding.bias']
    ref = {k: sd[k].clone().requires_grad_(True) for k in names}
    ropt = torch.optim.Adam([ref[k] for k in names], lr=2e-3, amsgrad=True)
    assert [tuple(p.shape) for p in model.get_parameters()] == [tuple(ref[k].shape) for k in names]
    xg = x.cuda()
    ours, theirs = [], []
    for _ in range(15):
        opt.zero_grad()
        ours.append(model.loss({'images': xg[None]}, chunk_size=40)['loss'])
        opt.step()
        ropt.zero_grad()
        lo, go = co.linear_ae_loss({k: v.detach() for k, v in ref.items()}, hp, x, None, chunk_size=40)
        for k in names:
            ref[k].grad = go[k]
        ropt.step()
        theirs.append(lo['loss'])
    ours, theirs = np.array(ours), np.array(theirs)
    assert theirs[-1] < 0.9 * theirs[0], theirs
    assert np.abs(ours - theirs).max() <= 1e-4 * theirs[0], (ours, theirs)
