"""CPU tests of the ARHMM oracle (parity unpinned: ssm is absent, see oracle/arhmm_oracle.py).
The restatement is validated by brute force, scipy, invariants and EM monotonicity, and frozen by
a committed regression fixture."""

import numpy as np
import pytest
from scipy.stats import multivariate_normal

from oracle import arhmm_oracle as ao
from tests.helpers import load_golden


@pytest.mark.parametrize('K,T,lags', [(3, 6, 1), (2, 8, 2), (4, 5, 0)])
def test_messages_and_viterbi_match_brute_force(K, T, lags):
    p = ao.synth_params(K, 2, lags, seed=K * 7 + T, mix=0.2)
    x = ao.sample(p, T, np.random.RandomState(T))[1]
    ll = ao.ar_log_likelihoods(x, p.As, p.bs, p.Sigmas, p.lags)
    g, xi, lz = ao.expected_states(p.log_pi0, p.log_Ps, ll)
    gb, xib, lzb, path = ao.brute_force(p.log_pi0, p.log_Ps, ll)
    assert np.abs(g - gb).max() < 1e-12
    assert np.abs(xi - xib).max() < 1e-12
    assert abs(lz - lzb) < 1e-10
    assert np.array_equal(ao.viterbi(p.log_pi0, p.log_Ps, ll), path)
    assert abs(ao.log_normalizer(p.log_pi0, p.log_Ps, ll) - lzb) < 1e-10


def test_emissions_match_scipy():
    p = ao.synth_params(3, 4, 2, seed=1)
    x = ao.sample(p, 30, np.random.RandomState(0))[1]
    ll = ao.ar_log_likelihoods(x, p.As, p.bs, p.Sigmas, p.lags)
    for k in range(3):
        for t in range(30):
            if t < 2:
                ref = multivariate_normal.logpdf(x[t], np.zeros(4), np.eye(4))
            else:
                mu = p.As[k][:, :4] @ x[t - 1] + p.As[k][:, 4:] @ x[t - 2] + p.bs[k]
                ref = multivariate_normal.logpdf(x[t], mu, p.Sigmas[k])
            assert abs(ll[t, k] - ref) < 1e-9


def test_posterior_invariants():
    p = ao.synth_params(6, 5, 2, seed=2)
    x = ao.sample(p, 200, np.random.RandomState(1))[1]
    g, xi, lz = ao.e_step(p, [x])[0]
    np.testing.assert_allclose(g.sum(1), 1, atol=1e-12)
    np.testing.assert_allclose(xi.sum(0), g[1:].sum(0), atol=1e-9)    # sum_j xi_t(j,k) = gamma_{t+1}(k)
    np.testing.assert_allclose(xi.sum(1), g[:-1].sum(0), atol=1e-9)   # sum_k xi_t(j,k) = gamma_t(j)
    np.testing.assert_allclose(xi.sum(), 199, atol=1e-9)


def test_golden_regression_fixture():
    g = load_golden('arhmm_k4_d3_l2')
    p = ao.ARHMMParams(g['log_pi0'], g['log_Ps'], g['As'], g['bs'], g['Sigmas'], int(g['lags']))
    off = np.concatenate([[0], np.cumsum(g['lengths'])])
    for i in range(len(g['lengths'])):
        x = g['x'][off[i]:off[i + 1]]
        ll = ao.ar_log_likelihoods(x, p.As, p.bs, p.Sigmas, p.lags)
        np.testing.assert_allclose(ll, g['ll'][off[i]:off[i + 1]], rtol=1e-10, atol=1e-10)
        ez, ezz, lz = ao.expected_states(p.log_pi0, p.log_Ps, ll)
        np.testing.assert_allclose(ez, g['Ez'][off[i]:off[i + 1]], atol=1e-12)
        np.testing.assert_allclose(ezz, g['Ezz'][i], atol=1e-10)
        assert abs(lz - g['logZ'][i]) < 1e-9
        assert np.array_equal(ao.viterbi(p.log_pi0, p.log_Ps, ll), g['z'][off[i]:off[i + 1]])


@pytest.mark.parametrize('transitions', ['stationary', 'sticky'])
def test_em_is_monotone(transitions):
    truth = ao.synth_params(3, 2, 1, seed=4, mix=0.3)
    rng = np.random.RandomState(3)
    xs = [ao.sample(truth, 150, rng)[1] for _ in range(4)]
    p = ao.synth_params(3, 2, 1, seed=9, mix=0.3)
    prev = ao.log_likelihood(p, xs)
    for _ in range(5):
        p = ao.m_step(p, xs, ao.e_step(p, xs), transitions=transitions, kappa=0.0)
        cur = ao.log_likelihood(p, xs)
        assert cur >= prev - 1e-6 * abs(prev)
        prev = cur


def test_log_likelihood_accepts_single_array():
    p = ao.synth_params(3, 2, 1, seed=4)
    x = ao.sample(p, 50, np.random.RandomState(0))[1]
    assert abs(ao.log_likelihood(p, x) - ao.log_likelihood(p, [x])) < 1e-12
