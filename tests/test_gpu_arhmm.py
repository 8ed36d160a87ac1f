"""GPU parity tests of hot path 2 (ARHMM E-step / log-likelihood / Viterbi) through
``behavenet_b200.ssm.HMM`` (which calls the C ABI), against the fp64 CPU oracle.

Tolerances (north_star): posteriors within 1e-5 absolute, Viterbi paths bit-exact.  Pairwise sums
sum_t xi_t are checked at 1e-5 * T (they are sums of T posteriors); log normalisers at 1e-6 relative.
"""

import numpy as np
import pytest

from oracle import arhmm_oracle as ao
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def make_hmm(p, transitions='stationary'):
    from behavenet_b200.ssm import HMM
    obs = 'ar' if p.lags > 0 else 'gaussian'
    kw = {'lags': p.lags} if p.lags > 0 else None
    hmm = HMM(p.K, p.D, observations=obs, observation_kwargs=kw, transitions=transitions)
    hmm.init_state_distn.log_pi0 = p.log_pi0.copy()
    hmm.transitions.log_Ps = p.log_Ps.copy()
    hmm.observations.As = p.As.copy()
    hmm.observations.bs = p.bs.copy()
    hmm.observations.Sigmas = p.Sigmas.copy()
    return hmm


def check_against_oracle(p, xs, post_tol=1e-5):
    hmm = make_hmm(p)
    ref = ao.e_step(p, xs)
    tot = 0.0
    for x, (g, j, lz) in zip(xs, ref):
        Ez, Ezz, logZ = hmm.expected_states(x)
        assert Ez.shape == g.shape and Ezz.shape == (1,) + j.shape
        assert np.abs(Ez - g).max() < post_tol
        assert np.abs(Ezz[0] - j).max() < post_tol * max(1, x.shape[0])
        assert abs(logZ - lz) <= 1e-6 * abs(lz) + 1e-4
        np.testing.assert_allclose(Ez.sum(1), 1.0, atol=1e-5)
        tot += lz
    ll = hmm.log_likelihood(xs)
    assert abs(ll - tot) <= 1e-6 * abs(tot) + 1e-4
    zs = hmm.most_likely_states_batch(xs)
    for x, z in zip(xs, zs):
        zr = ao.most_likely_states(p, x)
        assert z.dtype == np.int64 and np.array_equal(z, zr)          # bit-exact paths
    assert np.array_equal(hmm.most_likely_states(xs[0]), zs[0])


def test_golden_fixture_ragged_trials():
    """Committed fixture (K=4, D=3, lags=2) with ragged trials, including T=1 and T=3 <= lags+1."""
    g = load_golden('arhmm_k4_d3_l2')
    p = ao.ARHMMParams(g['log_pi0'], g['log_Ps'], g['As'], g['bs'], g['Sigmas'], int(g['lags']))
    off = np.concatenate([[0], np.cumsum(g['lengths'])])
    xs = [g['x'][off[i]:off[i + 1]] for i in range(len(g['lengths']))]
    hmm = make_hmm(p)
    for i, x in enumerate(xs):
        Ez, Ezz, logZ = hmm.expected_states(x)
        assert np.abs(Ez - g['Ez'][off[i]:off[i + 1]]).max() < 1e-5
        assert np.abs(Ezz[0] - g['Ezz'][i]).max() < 1e-5 * max(1, len(x))
        assert abs(logZ - g['logZ'][i]) < 1e-4
    zs = hmm.most_likely_states_batch(xs)
    assert np.array_equal(np.concatenate(zs), g['z'])
    check_against_oracle(p, xs)


@pytest.mark.parametrize('K,D,lags', [(16, 12, 2), (2, 1, 1), (5, 7, 3), (32, 4, 1), (3, 16, 0), (9, 30, 1)])
def test_estep_viterbi_loglik_shapes(K, D, lags):
    p = ao.synth_params(K, D, lags, seed=K + D, mix=0.05)
    rng = np.random.RandomState(1)
    if lags == 0:
        p.As = np.zeros((K, D, 0))
    xs = [ao.sample(p, T, rng)[1].astype(np.float32) for T in (150, 37, 64, 2)]
    check_against_oracle(p, xs)


def test_c4_geometry_subset_and_properties():
    """Config C4 shape (K=16, D=12, lag 2, T=1000): 24 trials against the oracle; then the full
    2048 x 1000 problem through size-independent properties (posterior rows sum to 1, the pairwise
    sums have row sums equal to the summed posteriors, path states in range, and the sharded
    log-likelihood is additive over trials)."""
    p = ao.synth_params()
    X = ao.sample_batch(p, 2048, 1000, seed=0)
    check_against_oracle(p, [X[i] for i in range(24)])
    import torch
    hmm = make_hmm(p)
    xs = [X[i] for i in range(X.shape[0])]
    st = hmm._stage(xs)
    Ez, Ezz, logZ = hmm._run_estep(st, True)
    torch.cuda.synchronize()
    assert float((Ez.sum(1) - 1).abs().max()) < 1e-5
    Ez3 = Ez.view(2048, 1000, 16)
    # sum_k xi_t(j,k) = gamma_t(j) for t < T-1
    assert float((Ezz.sum(2) - Ez3[:, :-1].sum(1)).abs().max()) < 2e-3
    assert float((Ezz.sum((1, 2)) - 999).abs().max()) < 1e-2
    ll_all = hmm.log_likelihood(xs)
    ll_half = hmm.log_likelihood(xs[:1024]) + hmm.log_likelihood(xs[1024:])
    assert abs(ll_all - ll_half) < 1e-6 * abs(ll_all)
    assert abs(ll_all - float(logZ.sum())) < 1e-9 * abs(ll_all)
    zs = hmm.most_likely_states_batch(xs[:256])
    z = np.concatenate(zs)
    assert z.min() >= 0 and z.max() < 16
    # the Viterbi path agrees with the posterior mode on confidently decoded steps
    g = Ez3[:256].reshape(-1, 16).cpu().numpy()
    conf = g.max(1) > 0.999
    assert (z[conf] == g.argmax(1)[conf]).mean() > 0.999


def test_em_iteration_matches_oracle_m_step_and_is_monotone():
    p = ao.synth_params(4, 3, 1, seed=11, mix=0.05)
    rng = np.random.RandomState(2)
    xs = [ao.sample(p, T, rng)[1].astype(np.float32) for T in (300, 200, 250)]
    q = ao.synth_params(4, 3, 1, seed=12, mix=0.3)          # start away from the truth
    hmm = make_hmm(q)
    lls = hmm.fit(xs, method='em', num_iters=1, initialize=False)
    ref = ao.m_step(q, xs, ao.e_step(q, xs))
    np.testing.assert_allclose(hmm.observations.As, ref.As, atol=2e-4)
    np.testing.assert_allclose(hmm.observations.bs, ref.bs, atol=2e-4)
    np.testing.assert_allclose(hmm.observations.Sigmas, ref.Sigmas, atol=2e-4)
    np.testing.assert_allclose(hmm.transitions.log_Ps, ref.log_Ps, atol=1e-3)
    np.testing.assert_allclose(hmm.init_state_distn.log_pi0, ref.log_pi0, atol=1e-4)
    more = hmm.fit(xs, method='em', num_iters=5, initialize=False)
    seq = lls + more[1:]
    assert all(b >= a - 1e-3 * abs(a) for a, b in zip(seq[:-1], seq[1:])), seq


def test_reference_call_sequence_runs():
    """The exact sequence of calls arhmm_grid_search.main makes (lines 131-209)."""
    import pickle
    from behavenet_b200.ssm import HMM
    np.random.seed(0)
    p = ao.synth_params(4, 6, 1, seed=5, mix=0.1)
    rng = np.random.RandomState(3)
    train = [ao.sample(p, 120, rng)[1].astype(np.float32) for _ in range(6)]
    hmm = HMM(4, 6, observations='ar', observation_kwargs={'lags': 1}, transitions='sticky',
              transition_kwargs={'kappa': 10.0})
    hmm.initialize(train)
    hmm.observations.initialize(train, localize=True)
    hmm.hparams = {'anything': 1}
    prev = None
    for epoch in range(3):
        if epoch > 0:
            hmm.fit(train, method='em', num_iters=1, initialize=False)
        ll = hmm.log_likelihood(train)
        assert np.isfinite(ll)
        if prev is not None:
            assert ll > prev - 1e-3 * abs(prev)
        prev = ll
    assert np.isfinite(hmm.log_likelihood(train[0]))          # single array (line 196)
    zs = [hmm.most_likely_states(x) for x in train]
    usage = np.bincount(np.concatenate(zs), minlength=hmm.K)
    hmm.permute(np.argsort(usage)[::-1])
    zs2 = [hmm.most_likely_states(x) for x in train]
    usage2 = np.bincount(np.concatenate(zs2), minlength=hmm.K)
    assert np.array_equal(usage2, np.sort(usage)[::-1])
    clone = pickle.loads(pickle.dumps(hmm))
    assert np.array_equal(clone.most_likely_states(train[0]), zs2[0])
    z, x = hmm.sample(50)
    assert z.shape == (50,) and x.shape == (50, 6)
    assert hmm.transitions.transition_matrix.shape == (4, 4)


@pytest.mark.parametrize('K,D,lags', [(16, 12, 2), (5, 7, 3), (32, 4, 1), (3, 16, 0)])
def test_cuda_core_and_tensor_core_emissions_agree(K, D, lags):
    """bn_set_tensor_core_mode(0) keeps the E-step on the fp32 CUDA-core emission kernel; mode 1 (default)
    uses the tcgen05 3xTF32 kernel.  Both must meet the oracle tolerance and agree with each other."""
    from behavenet_b200 import _lib
    p = ao.synth_params(K, D, lags, seed=3 * K + D, mix=0.05)
    if lags == 0:
        p.As = np.zeros((K, D, 0))
    rng = np.random.RandomState(4)
    xs = [ao.sample(p, T, rng)[1].astype(np.float32) for T in (300, 129, 5)]
    out = {}
    try:
        for mode in (0, 1):
            _lib.lib().bn_set_tensor_core_mode(mode)
            check_against_oracle(p, xs)
            hmm = make_hmm(p)
            out[mode] = [hmm.expected_states(x) for x in xs]
    finally:
        _lib.lib().bn_set_tensor_core_mode(1)
    for (g0, j0, l0), (g1, j1, l1) in zip(out[0], out[1]):
        assert np.abs(g0 - g1).max() < 1e-5
        assert abs(l0 - l1) <= 1e-6 * abs(l0) + 1e-4


def test_many_ragged_trials_batched():
    """600 ragged trials in ONE launch: every emission CTA walks several trials and time tiles, scan warps
    carry trials of different lengths.  Batched results must equal the per-trial oracle (subset) and the
    CUDA-core emission path (all trials)."""
    import torch
    from behavenet_b200 import _lib
    K, D, lags = 6, 4, 2
    p = ao.synth_params(K, D, lags, seed=21, mix=0.05)
    rng = np.random.RandomState(5)
    lens = rng.randint(1, 400, size=600)
    lens[:4] = [1, 2, 3, 399]
    xs = [ao.sample(p, int(T), rng)[1].astype(np.float32) for T in lens]
    hmm = make_hmm(p)
    st = hmm._stage(xs)
    out = {}
    try:
        for mode in (0, 1):
            _lib.lib().bn_set_tensor_core_mode(mode)
            Ez, Ezz, logZ = hmm._run_estep(st, True)
            torch.cuda.synchronize()
            out[mode] = (Ez.cpu().numpy(), Ezz.cpu().numpy(), logZ.cpu().numpy())
    finally:
        _lib.lib().bn_set_tensor_core_mode(1)
    assert np.abs(out[0][0] - out[1][0]).max() < 1e-5
    assert np.abs(out[0][2] - out[1][2]).max() < 1e-3
    off = np.concatenate([[0], np.cumsum(lens)])
    for i in list(range(8)) + [77, 300, 599]:
        g, j, lz = ao.e_step(p, [xs[i]])[0]
        for mode in (0, 1):
            assert np.abs(out[mode][0][off[i]:off[i + 1]] - g).max() < 1e-5
            assert np.abs(out[mode][1][i] - j).max() < 1e-5 * max(1, lens[i])
            assert abs(out[mode][2][i] - lz) <= 1e-6 * abs(lz) + 1e-4


def test_long_trials_stay_normalised():
    """5000-step trials: any drift of the message scale (a lagged / marginally stable normalisation)
    would overflow fp32 long before the end."""
    p = ao.synth_params(8, 6, 2, seed=9, mix=0.05)
    rng = np.random.RandomState(6)
    xs = [ao.sample(p, T, rng)[1].astype(np.float32) for T in (5000, 4097)]
    check_against_oracle(p, xs)


@pytest.mark.parametrize('env', [{'BN_SCAN': '4'}, {'BN_EMIT': '1'}, {'BN_SCAN': '1'}])
def test_alternative_estep_kernels_match_oracle(env):
    """The kernels behind the diagnostic switches (SPL-states-per-lane scan, one-stage tcgen05 emission kernel,
    three-pass scan) stay parity-green: the switches are read once per process, so each runs in its own."""
    import os
    import subprocess
    import sys
    from tests.helpers import ROOT
    code = (
        "import numpy as np\n"
        "from oracle import arhmm_oracle as ao\n"
        "from tests.test_gpu_arhmm import check_against_oracle\n"
        "for K, D, lags in [(16, 12, 2), (32, 4, 1), (8, 8, 1)]:\n"
        "    p = ao.synth_params(K, D, lags, seed=K + D, mix=0.05)\n"
        "    X = ao.sample_batch(p, 6, 300, seed=1)\n"
        "    xs = [X[i][:n] for i, n in enumerate([300, 1, 2, 17, 129, 257])]\n"
        "    check_against_oracle(p, xs)\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, env=dict(os.environ, **env), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stdout[-2000:] + r.stderr[-4000:]
