"""CPU oracle for hot path 2 (ARHMM E-step / log-likelihood / Viterbi).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference delegates this path to the third-party ``ssm`` package
(github.com/slinderman/ssm, branch ``behavenet-no-cython`` -- pinned by branch name only,
/root/reference/docs/source/installation.rst:62-69); it is neither vendored under /root/reference
nor installed in this image, and the reference's own tests only ever use a mock HMM
(/root/reference/tests/test_plotting/test_arhmm_utils.py:30-35).  This file therefore restates the
published ssm algorithm (fp64, log-space messages, numba loops) and is
anchored on the reference's call sites:
    fitting/arhmm_grid_search.py:132-137 (construct/initialise)   :170 (fit, method='em')
    :173-196 (log_likelihood)   :201-204 (most_likely_states, permute)   fitting/eval.py:167
It is validated by internal consistency instead (tests/test_oracle_arhmm.py): brute-force
enumeration of all K^T paths, scipy's multivariate-normal logpdf, posterior normalisation, and EM
monotonicity.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this module; the product never does.
"""

import itertools
import math

import numpy as np

try:                                    # numba is ssm's own execution model for these loops
    import numba
    _jit = numba.njit(cache=False)
except Exception:                       # pragma: no cover - pure-python loops still work
    numba = None

    def _jit(f):
        return f

LN2PI = math.log(2 * math.pi)


# ------------------------------------------------------------------------------------------------
# emissions: ssm AutoRegressiveObservations.log_likelihoods
# ------------------------------------------------------------------------------------------------

def ar_means(x, As, bs, lags):
    """(T-L, K, D) AR means for t >= L; block l of ``As`` multiplies x_{t-l-1}."""
    T, D = x.shape
    K = As.shape[0]
    mus = np.tile(bs[None, :, :], (T - lags, 1, 1)).astype(np.float64)
    for l in range(lags):
        hist = x[lags - l - 1:T - l - 1]                              # x_{t-l-1}, t = L..T-1
        mus += np.einsum('td,ked->tke', hist, As[:, :, l * D:(l + 1) * D])
    return mus


def mvn_logpdf(resid, Sigma):
    """log N(resid; 0, Sigma) rows, via Cholesky (ssm.stats.multivariate_normal_logpdf)."""
    D = resid.shape[-1]
    L = np.linalg.cholesky(Sigma)
    y = np.linalg.solve(L, resid.T).T if resid.shape[0] else resid
    return -0.5 * D * LN2PI - np.log(np.diag(L)).sum() - 0.5 * (y ** 2).sum(-1)


def ar_log_likelihoods(x, As, bs, Sigmas, lags, mu_init=None, Sigmas_init=None):
    """(T, K) emission log-likelihoods.  First ``lags`` steps: N(mu_init_k, Sigma_init_k)
    (zeros / identity in ssm); afterwards N(sum_l A_k^(l) x_{t-l-1} + b_k, Sigma_k)."""
    x = np.asarray(x, dtype=np.float64)
    T, D = x.shape
    K = As.shape[0]
    mu_init = np.zeros((K, D)) if mu_init is None else mu_init
    Sigmas_init = np.tile(np.eye(D)[None], (K, 1, 1)) if Sigmas_init is None else Sigmas_init
    ll = np.empty((T, K))
    n0 = min(lags, T)
    for k in range(K):
        ll[:n0, k] = mvn_logpdf(x[:n0] - mu_init[k], Sigmas_init[k])
    if T > lags:
        mus = ar_means(x, As, bs, lags)
        for k in range(K):
            ll[lags:, k] = mvn_logpdf(x[lags:] - mus[:, k], Sigmas[k])
    return ll


# ------------------------------------------------------------------------------------------------
# message passing: ssm.messages (numba) -- log-space, fp64
# ------------------------------------------------------------------------------------------------

@_jit
def _logsumexp(v):
    m = -np.inf
    for i in range(v.shape[0]):
        if v[i] > m:
            m = v[i]
    if m == -np.inf:
        return m
    s = 0.0
    for i in range(v.shape[0]):
        s += math.exp(v[i] - m)
    return m + math.log(s)


@_jit
def _forward(log_pi0, log_P, ll, alphas):
    T, K = ll.shape
    tmp = np.empty(K)
    for k in range(K):
        alphas[0, k] = log_pi0[k] + ll[0, k]
    for t in range(T - 1):
        for k in range(K):
            for j in range(K):
                tmp[j] = alphas[t, j] + log_P[j, k]
            alphas[t + 1, k] = _logsumexp(tmp) + ll[t + 1, k]
    return _logsumexp(alphas[T - 1])


@_jit
def _backward(log_P, ll, betas):
    T, K = ll.shape
    tmp = np.empty(K)
    for k in range(K):
        betas[T - 1, k] = 0.0
    for t in range(T - 2, -1, -1):
        for j in range(K):
            for k in range(K):
                tmp[k] = log_P[j, k] + ll[t + 1, k] + betas[t + 1, k]
            betas[t, j] = _logsumexp(tmp)


@_jit
def _joints_sum(alphas, betas, ll, log_P, out):
    """sum_t xi_t, each xi_t normalised on its own (ssm _compute_stationary_expected_joints)."""
    T, K = ll.shape
    tmp = np.empty((K, K))
    for t in range(T - 1):
        m = -np.inf
        for j in range(K):
            for k in range(K):
                v = alphas[t, j] + log_P[j, k] + ll[t + 1, k] + betas[t + 1, k]
                tmp[j, k] = v
                if v > m:
                    m = v
        s = 0.0
        for j in range(K):
            for k in range(K):
                tmp[j, k] = math.exp(tmp[j, k] - m)
                s += tmp[j, k]
        for j in range(K):
            for k in range(K):
                out[j, k] += tmp[j, k] / s


@_jit
def _viterbi(log_pi0, log_P, ll, z):
    """Backward max-sum, first-index argmax on ties (ssm.messages.viterbi)."""
    T, K = ll.shape
    scores = np.zeros((T, K))
    args = np.zeros((T, K), np.int64)
    for t in range(T - 2, -1, -1):
        for j in range(K):
            best = -np.inf
            arg = 0
            for k in range(K):
                v = log_P[j, k] + scores[t + 1, k] + ll[t + 1, k]
                if v > best:
                    best = v
                    arg = k
            args[t + 1, j] = arg
            scores[t, j] = best
    best = -np.inf
    arg = 0
    for j in range(K):
        v = scores[0, j] + log_pi0[j] + ll[0, j]
        if v > best:
            best = v
            arg = j
    z[0] = arg
    for t in range(1, T):
        z[t] = args[t, z[t - 1]]


def expected_states(log_pi0, log_Ps, ll):
    """ssm.HMM.expected_states for stationary transitions.

    Returns (Ez (T,K), sum_t Ezz (K,K), log normaliser)."""
    ll = np.ascontiguousarray(ll, dtype=np.float64)
    T, K = ll.shape
    alphas = np.empty((T, K))
    betas = np.empty((T, K))
    logZ = _forward(log_pi0, log_Ps, ll, alphas)
    _backward(log_Ps, ll, betas)
    g = alphas + betas
    g -= g.max(1, keepdims=True)
    g = np.exp(g)
    g /= g.sum(1, keepdims=True)
    joints = np.zeros((K, K))
    _joints_sum(alphas, betas, ll, log_Ps, joints)
    return g, joints, logZ


def log_normalizer(log_pi0, log_Ps, ll):
    ll = np.ascontiguousarray(ll, dtype=np.float64)
    return _forward(log_pi0, log_Ps, ll, np.empty(ll.shape))


def viterbi(log_pi0, log_Ps, ll):
    ll = np.ascontiguousarray(ll, dtype=np.float64)
    z = np.zeros(ll.shape[0], np.int64)
    _viterbi(log_pi0, log_Ps, ll, z)
    return z


# ------------------------------------------------------------------------------------------------
# model-level wrappers (what the reference calls on the ssm.HMM object)
# ------------------------------------------------------------------------------------------------

class ARHMMParams:
    """Plain container with ssm's parameterisation (fp64)."""

    def __init__(self, log_pi0, log_Ps, As, bs, Sigmas, lags):
        self.log_pi0 = np.asarray(log_pi0, np.float64)
        self.log_Ps = np.asarray(log_Ps, np.float64)
        self.As = np.asarray(As, np.float64)
        self.bs = np.asarray(bs, np.float64)
        self.Sigmas = np.asarray(Sigmas, np.float64)
        self.lags = int(lags)
        self.K, self.D = self.bs.shape


def synth_params(K=16, D=12, lags=2, seed=0, mix=0.04):
    """ssm-style parameters for config C4: A_k = 0.95 [R_k | 0] + noise with
    R_k rotations near a common one (``mix`` sets how distinguishable the states are, so that
    posteriors are not degenerate), b ~ 0.1 N(0,1), Sigma_k = 0.1 I + 0.01 G G^T,
    P = 0.95 I + 0.05 U (rows normalised), pi0 uniform."""
    rng = np.random.RandomState(seed)
    q0, _ = np.linalg.qr(rng.randn(D, D))
    As = np.zeros((K, D, D * lags))
    for k in range(K):
        q, _ = np.linalg.qr(q0 + mix * rng.randn(D, D))
        if lags > 0:
            As[k, :, :D] = 0.95 * q
            As[k] += 0.01 * rng.randn(D, D * lags)
    bs = 0.1 * rng.randn(K, D)
    G = rng.randn(K, D, D)
    Sigmas = 0.1 * np.eye(D)[None] + 0.01 * np.einsum('kij,klj->kil', G, G)
    P = 0.95 * np.eye(K) + 0.05 * rng.rand(K, K)
    P /= P.sum(1, keepdims=True)
    return ARHMMParams(-np.log(K) * np.ones(K), np.log(P), As, bs, Sigmas, lags)


def sample(params, T, rng):
    """Draw (z, x) from the ARHMM (ssm.HMM.sample semantics: first ``lags`` obs ~ N(0, I))."""
    K, D, L = params.K, params.D, params.lags
    P = np.exp(params.log_Ps)
    chol = np.linalg.cholesky(params.Sigmas)
    z = np.zeros(T, np.int64)
    x = np.zeros((T, D))
    z[0] = rng.choice(K, p=np.exp(params.log_pi0))
    for t in range(T):
        if t > 0:
            z[t] = rng.choice(K, p=P[z[t - 1]])
        if t < L:
            x[t] = rng.randn(D)
        else:
            hist = np.concatenate([x[t - l - 1] for l in range(L)]) if L else np.zeros(0)
            x[t] = params.As[z[t]] @ hist + params.bs[z[t]] + chol[z[t]] @ rng.randn(D)
    return z, x


def sample_batch(params, n_trials, T, seed=0, dtype=np.float32):
    """Vectorised-over-trials sampler for benchmark-size inputs: (n_trials, T, D)."""
    rng = np.random.RandomState(seed)
    K, D, L = params.K, params.D, params.lags
    cdf = np.cumsum(np.exp(params.log_Ps), 1)
    chol = np.linalg.cholesky(params.Sigmas)
    z = rng.randint(0, K, size=n_trials)
    x = np.zeros((n_trials, T, D))
    for t in range(T):
        if t > 0:
            u = rng.rand(n_trials)
            z = np.minimum((u[:, None] > cdf[z]).sum(1), K - 1)
        noise = rng.randn(n_trials, D)
        if t < L:
            x[:, t] = noise
        else:
            hist = (np.concatenate([x[:, t - l - 1] for l in range(L)], 1) if L
                    else np.zeros((n_trials, 0)))
            x[:, t] = (np.einsum('nde,ne->nd', params.As[z], hist) + params.bs[z]
                       + np.einsum('nde,ne->nd', chol[z], noise))
    return x.astype(dtype)


def e_step(params, datas):
    """[ (Ez, sum Ezz, logZ) for each trial ] -- the list comprehension inside ssm's fit."""
    out = []
    for x in datas:
        ll = ar_log_likelihoods(x, params.As, params.bs, params.Sigmas, params.lags)
        out.append(expected_states(params.log_pi0, params.log_Ps, ll))
    return out


def log_likelihood(params, datas):
    """ssm.HMM.log_likelihood: accepts one (T,D) array or a list (arhmm_grid_search.py:173,196)."""
    if isinstance(datas, np.ndarray) and datas.ndim == 2:
        datas = [datas]
    tot = 0.0
    for x in datas:
        ll = ar_log_likelihoods(x, params.As, params.bs, params.Sigmas, params.lags)
        tot += log_normalizer(params.log_pi0, params.log_Ps, ll)
    return tot


def most_likely_states(params, x):
    ll = ar_log_likelihoods(x, params.As, params.bs, params.Sigmas, params.lags)
    return viterbi(params.log_pi0, params.log_Ps, ll)


# ------------------------------------------------------------------------------------------------
# M-step (ssm InitialStateDistribution / StationaryTransitions /
# StickyTransitions / AutoRegressiveObservations m_step)
# ------------------------------------------------------------------------------------------------

def ar_sufficient_stats(params, datas, expectations):
    """Per-state weighted regression statistics of x_t on phi_t = [x_{t-1..t-L}, 1], t >= L."""
    K, D, L = params.K, params.D, params.lags
    P = D * L + 1
    Sxx = np.zeros((K, P, P))
    Sxy = np.zeros((K, P, D))
    Syy = np.zeros((K, D, D))
    Sn = np.zeros(K)
    for x, (Ez, _, _) in zip(datas, expectations):
        x = np.asarray(x, np.float64)
        T = x.shape[0]
        if T <= L:
            continue
        phi = np.concatenate([x[L - l - 1:T - l - 1] for l in range(L)] + [np.ones((T - L, 1))], 1)
        y = x[L:]
        w = Ez[L:]
        Sxx += np.einsum('tk,ti,tj->kij', w, phi, phi)
        Sxy += np.einsum('tk,ti,tj->kij', w, phi, y)
        Syy += np.einsum('tk,ti,tj->kij', w, y, y)
        Sn += w.sum(0)
    return Sxx, Sxy, Syy, Sn


def m_step(params, datas, expectations, transitions='stationary', kappa=0.0, alpha=1.0,
           l2_penalty=1e-8, nu0=1e-4, psi0=1e-4):
    """One M-step; returns new ARHMMParams."""
    K, D, L = params.K, params.D, params.lags
    pi0 = sum(Ez[0] for Ez, _, _ in expectations) + 1e-8
    log_pi0 = np.log(pi0 / pi0.sum())
    J = sum(Ezz for _, Ezz, _ in expectations)
    if transitions == 'sticky':
        J = J + 1e-16 + kappa * np.eye(K) + (alpha - 1.0)
        Pm = J / J.sum(1, keepdims=True) + 1e-16
        log_Ps = np.log(Pm)
    else:
        Pm = J + 1e-32
        Pm = Pm / Pm.sum(1, keepdims=True)
        log_Ps = np.log(Pm)
        log_Ps = log_Ps - np.log(np.exp(log_Ps).sum(1, keepdims=True))
    Sxx, Sxy, Syy, Sn = ar_sufficient_stats(params, datas, expectations)
    As = np.zeros_like(params.As)
    bs = np.zeros_like(params.bs)
    Sigmas = np.zeros_like(params.Sigmas)
    J0 = l2_penalty * np.eye(D * L + 1)
    for k in range(K):
        W = np.linalg.solve(Sxx[k] + J0, Sxy[k]).T                    # (D, D*L+1)
        As[k], bs[k] = W[:, :D * L], W[:, -1]
        WSxy = W @ Sxy[k]
        sqerr = Syy[k] - WSxy.T - WSxy + W @ Sxx[k] @ W.T
        Sigmas[k] = (sqerr + psi0 * np.eye(D)) / (nu0 + Sn[k] + D + 1)
        Sigmas[k] = 0.5 * (Sigmas[k] + Sigmas[k].T)
    return ARHMMParams(log_pi0, log_Ps, As, bs, Sigmas, L)


# ------------------------------------------------------------------------------------------------
# brute force (validation of the restatement itself)
# ------------------------------------------------------------------------------------------------

def brute_force(log_pi0, log_Ps, ll):
    """Enumerate all K^T paths: (gamma (T,K), sum_t xi (K,K), logZ, argmax path)."""
    T, K = ll.shape
    paths = list(itertools.product(range(K), repeat=T))
    lp = np.empty(len(paths))
    for i, p in enumerate(paths):
        v = log_pi0[p[0]] + ll[0, p[0]]
        for t in range(1, T):
            v += log_Ps[p[t - 1], p[t]] + ll[t, p[t]]
        lp[i] = v
    m = lp.max()
    w = np.exp(lp - m)
    logZ = m + math.log(w.sum())
    w /= w.sum()
    gamma = np.zeros((T, K))
    xi = np.zeros((K, K))
    for wi, p in zip(w, paths):
        for t in range(T):
            gamma[t, p[t]] += wi
        for t in range(T - 1):
            xi[p[t], p[t + 1]] += wi
    return gamma, xi, logZ, np.array(paths[int(np.argmax(lp))])
