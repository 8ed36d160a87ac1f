"""behavenet_b200.fitting.losses against the known-answer cases of the reference's tests/test_fitting/test_losses.py
(re-stated; /root/reference is not read at run time) and against the oracle's independent restatement."""

import numpy as np
import torch

from behavenet_b200.fitting import losses
from oracle import cae_oracle as co

LN2PI = np.log(2 * np.pi)


def test_mse_and_gaussian_ll_closed_forms():
    x = torch.rand((5, 3))
    assert losses.mse(x, x) == 0
    a = torch.tensor([1, 2, 3, 4, 5, 6], dtype=torch.float)
    b = torch.tensor([2, 3, 4, 5, 6, 7], dtype=torch.float)
    m = torch.tensor([1, 0, 1, 0, 1, 0], dtype=torch.float)
    assert losses.mse(a, b, m) == 0.5
    for std in (1, 2.0):
        const = -(0.5 * LN2PI + 0.5 * np.log(std ** 2)) * 3
        assert np.isclose(float(losses.gaussian_ll(x, x, masks=None, std=std)), const)
        ones, zeros, mk = torch.ones(5, 3), torch.zeros(5, 3), torch.zeros(5, 3)
        mk[:, 0] = 1
        assert np.isclose(float(losses.gaussian_ll(ones, zeros, masks=mk, std=std)), const - 0.5 / std ** 2)
        ll = losses.gaussian_ll(ones, zeros, std=std)
        mse_ = 2 * (-ll - (0.5 * LN2PI + 0.5 * np.log(std ** 2)) * 3) / 3 * std ** 2
        assert np.allclose(losses.gaussian_ll_to_mse(ll.numpy(), 3, gaussian_std=std, mse_std=1), mse_.numpy())
    assert losses.kl_div_to_std_normal(torch.zeros(1, 1), torch.zeros(1, 1)) == 0


def test_decomposed_kl_terms_agree_with_single_estimators_and_oracle():
    g = torch.Generator().manual_seed(0)
    z, mu, lv = torch.rand(5, 3, generator=g), torch.rand(5, 3, generator=g), torch.rand(5, 3, generator=g)
    ic, tc, dw = losses.decomposed_kl(z, mu, lv)
    assert losses.index_code_mi(z, mu, lv).item() == ic.item()
    assert losses.total_correlation(z, mu, lv).item() == tc.item()
    assert losses.dimension_wise_kl_to_std_normal(z, mu, lv).item() == dw.item()
    for a, b in zip((ic, tc, dw), co.decomposed_kl(z, mu, lv)):
        assert abs(float(a) - float(b)) < 1e-6
    assert abs(float(losses.kl_div_to_std_normal(mu, lv)) - float(co.kl_div_to_std_normal(mu, lv))) < 1e-6


def test_subspace_overlap():
    from scipy.linalg import null_space, orth
    A = torch.tensor([[1, 0, 0], [0, 1, 0]]).float()
    B = torch.tensor([[0, 0, 1]]).float()
    assert losses.subspace_overlap(A, B) == 0
    M = orth(np.random.RandomState(0).randn(15, 10)).T
    N = null_space(M)
    assert np.isclose(float(losses.subspace_overlap(torch.from_numpy(M), torch.from_numpy(N.T))), 0)
    k = 10
    eye = torch.from_numpy(np.eye(k)).float()
    assert losses.subspace_overlap(eye, eye) == 2 * k / ((2 * k) ** 2)
    assert np.isclose(float(losses.subspace_overlap(A, B, C=A)), float(torch.mean(
        (torch.cat([A, B, A]) @ torch.cat([A, B, A]).T - torch.eye(5)).pow(2))))


def test_triplet_loss_known_answers():
    tl = torch.nn.TripletMarginLoss(margin=1.0, p=2)
    n_batch, n_dims = 6, 3
    ds2 = np.concatenate([np.zeros(n_batch), np.ones(n_batch)])
    x = torch.zeros((n_batch, n_dims))
    for scale in (1.0, 2.0):                                     # far enough apart: every hinge is inactive
        assert np.isclose(losses.triplet_loss(tl, torch.cat([x, scale * torch.ones_like(x)], 0), ds2).item(), 0, atol=1e-5)
    t1 = 0.5
    loss = losses.triplet_loss(tl, torch.cat([x, t1 * torch.ones_like(x)], 0), ds2)
    assert np.isclose(loss.item(), (-np.sqrt(n_dims * t1 ** 2) + 1) * 2 / 3, atol=1e-5)
    t1, t2 = 0.25, 0.5
    ds3 = np.concatenate([np.zeros(n_batch), np.ones(n_batch), 2 * np.ones(n_batch)])
    loss = losses.triplet_loss(tl, torch.cat([x, t1 * torch.ones_like(x), t2 * torch.ones_like(x)], 0), ds3)
    v1, v2 = -np.sqrt(n_dims * t1 ** 2) + 1, -np.sqrt(n_dims * t2 ** 2) + 1
    assert np.isclose(loss.item(), (4 * v1 + 2 * v2) / 6, atol=1e-5)
    n_batch = 9
    x = torch.zeros((n_batch, n_dims))
    t1, t2, t3 = 0.1, 0.2, 0.3
    ds4 = np.concatenate([i * np.ones(n_batch) for i in range(4)])
    loss = losses.triplet_loss(tl, torch.cat([x] + [t * torch.ones_like(x) for t in (t1, t2, t3)], 0), ds4)
    v1, v2, v3 = (-np.sqrt(n_dims * t ** 2) + 1 for t in (t1, t2, t3))
    assert np.isclose(loss.item(), (6 * v1 + 4 * v2 + 2 * v3) / 12, atol=1e-5)


def test_r2_variance_weighted_matches_sklearn():
    """label_r2 of the PS-VAE / MSP losses (reference vaes.py:709-718 calls sklearn's r2_score per loss call):
    the validation-free restatement returns sklearn's value, including constant and perfectly predicted columns."""
    from sklearn.metrics import r2_score
    from behavenet_b200.fitting.losses import r2_variance_weighted
    rng = np.random.default_rng(0)
    for trial in range(60):
        n, m = int(rng.integers(2, 300)), int(rng.integers(1, 6))
        y = rng.standard_normal((n, m)).astype(np.float32)
        yh = y + rng.standard_normal((n, m)).astype(np.float32) * rng.random()
        if trial % 7 == 0:
            y[:, 0] = 1.5                       # constant target column: zero denominator
        if trial % 11 == 0:
            yh[:, 0] = y[:, 0]                  # perfect prediction: zero numerator
        if trial % 13 == 0:
            y[:] = 2.0                          # every denominator zero: uniform average
        a = r2_score(y, yh, multioutput='variance_weighted')
        b = r2_variance_weighted(y, yh)
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (trial, a, b)
