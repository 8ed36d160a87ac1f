"""ARHMM object with the subset of the ``ssm.HMM`` protocol that BehaveNet uses, on sm_100a kernels.

The reference constructs ``ssm.HMM(K, D, observations='ar', observation_kwargs={'lags': L},
transitions='stationary'|'sticky', transition_kwargs=...)`` and calls ``initialize``,
``observations.initialize``, ``fit(method='em', num_iters=1, initialize=False)``,
``log_likelihood``, ``most_likely_states``, ``permute``, ``sample`` and reads
``transitions.transition_matrix`` / ``observations.{As, bs, Sigmas, lags}``
(reference ``behavenet/fitting/arhmm_grid_search.py:132-209``, ``fitting/eval.py:167``,
``plotting/arhmm_utils.py:193-250, 845-975``).  This class provides exactly that surface.

Division of labour:
  * E-step (posteriors, pairwise sums, log normalisers), log-likelihood, Viterbi and the AR
    sufficient statistics run in ``libbehavenet_b200.so`` (``bn_arhmm_*``) -- no CPU path.
  * Parameters live on the host in fp64 numpy arrays with ssm's parameterisation, so the object
    pickles like the reference's (``pickle.dump(hmm)``, arhmm_grid_search.py:207-209).
  * The M-step solves (K small dense systems), k-means initialisation and sampling are host-side
    numpy, as in ssm; they consume the GPU-reduced statistics.
"""

import os

import numpy as np

from behavenet_b200 import _lib, parallel

__all__ = ['HMM']


def _logsumexp(a, axis=None, keepdims=False):
    m = np.max(a, axis=axis, keepdims=True)
    out = m + np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True))
    return out if keepdims else np.squeeze(out, axis=axis)


class InitialStateDistribution:
    def __init__(self, K):
        self.K = K
        self.log_pi0 = -np.log(K) * np.ones(K)

    @property
    def initial_state_distn(self):
        return np.exp(self.log_pi0 - _logsumexp(self.log_pi0))

    def permute(self, perm):
        self.log_pi0 = self.log_pi0[perm]

    def m_step(self, gamma0_sum):
        pi0 = gamma0_sum + 1e-8
        self.log_pi0 = np.log(pi0 / pi0.sum())


class StationaryTransitions:
    """Row-stochastic transition matrix; ssm initialisation 0.95 I + 0.05 U (rows normalised)."""

    def __init__(self, K):
        self.K = K
        Ps = 0.95 * np.eye(K) + 0.05 * np.random.rand(K, K)
        Ps /= Ps.sum(axis=1, keepdims=True)
        self.log_Ps = np.log(Ps)

    @property
    def transition_matrix(self):
        return np.exp(self.log_Ps - _logsumexp(self.log_Ps, axis=1, keepdims=True))

    def permute(self, perm):
        self.log_Ps = self.log_Ps[np.ix_(perm, perm)]

    def log_prior(self):
        return 0.0

    def m_step(self, joints_sum):
        P = joints_sum + 1e-32
        P = np.nan_to_num(P / P.sum(axis=-1, keepdims=True))
        P = np.where(P.sum(axis=-1, keepdims=True) == 0, 1.0 / self.K, P)
        log_P = np.log(P)
        self.log_Ps = log_P - _logsumexp(log_P, axis=-1, keepdims=True)


class StickyTransitions(StationaryTransitions):
    """Dirichlet prior with extra mass ``kappa`` on self-transitions."""

    def __init__(self, K, alpha=1, kappa=100):
        super().__init__(K)
        self.alpha = alpha
        self.kappa = kappa

    def log_prior(self):
        K = self.K
        log_P = self.log_Ps - _logsumexp(self.log_Ps, axis=1, keepdims=True)
        conc = self.alpha * np.ones((K, K)) + self.kappa * np.eye(K)
        return float(np.sum((conc - 1) * log_P))

    def m_step(self, joints_sum):
        J = joints_sum + 1e-16
        J = J + self.kappa * np.eye(self.K) + (self.alpha - 1) * np.ones((self.K, self.K))
        P = J / J.sum(axis=1, keepdims=True) + 1e-16
        if not np.all(P >= 0):
            raise ValueError('mode is well defined only for non-negative entries; check alpha >= 1')
        self.log_Ps = np.log(P)


class AutoRegressiveObservations:
    """x_t ~ N(sum_l A_k^(l) x_{t-l-1} + b_k, Sigma_k); first ``lags`` steps ~ N(0, I)."""

    def __init__(self, K, D, lags=1, l2_penalty_A=1e-8, l2_penalty_b=1e-8, nu0=1e-4, Psi0=1e-4):
        self.K, self.D, self.lags = K, D, lags
        self.As = np.zeros((K, D, D * lags))
        if lags > 0:
            self.As[:, :, :D] = 0.95 * np.array([_random_rotation(D) for _ in range(K)])
        self.bs = np.random.randn(K, D)
        self.Sigmas = np.tile(np.eye(D)[None], (K, 1, 1))
        self.l2_penalty_A, self.l2_penalty_b = l2_penalty_A, l2_penalty_b
        self.nu0, self.Psi0 = nu0, Psi0 * np.eye(D)

    def permute(self, perm):
        self.As, self.bs, self.Sigmas = self.As[perm], self.bs[perm], self.Sigmas[perm]

    def log_prior(self):
        return 0.0

    def initialize(self, datas, inputs=None, masks=None, tags=None, localize=True):
        """k-means clusters -> per-cluster linear regressions (ssm AutoRegressiveObservations
        .initialize with localize=True, called at arhmm_grid_search.py:137)."""
        from sklearn.cluster import KMeans
        from sklearn.linear_model import LinearRegression
        K, D, L = self.K, self.D, self.lags
        datas = [np.asarray(d, dtype=np.float64) for d in datas]
        datas = [d for d in datas if d.shape[0] > L]
        if L == 0:
            km = KMeans(K, n_init=10).fit(np.concatenate(datas))
            self.bs = km.cluster_centers_.copy()
            for k in range(K):
                pts = np.concatenate(datas)[km.labels_ == k]
                self.Sigmas[k] = np.cov(pts.T).reshape(D, D) + 1e-4 * np.eye(D) if len(pts) > D else np.eye(D)
            return
        if localize:
            km = KMeans(K, n_init=10).fit(np.concatenate(datas))
            zs = np.split(km.labels_, np.cumsum([len(d) for d in datas])[:-1])
            zs = [z[:-L] for z in zs]
        else:
            zs = [np.random.choice(K, size=len(d) - L) for d in datas]
        Sigmas = []
        for k in range(K):
            ts = [np.where(z == k)[0] for z in zs]
            Xs = [np.column_stack([d[t + l] for l in range(L - 1, -1, -1)]) for t, d in zip(ts, datas) if len(t)]
            ys = [d[t + L] for t, d in zip(ts, datas) if len(t)]
            if len(Xs) == 0 or sum(len(y) for y in ys) <= D * L + 1:
                Sigmas.append(np.eye(D))
                continue
            X, y = np.vstack(Xs), np.vstack(ys)
            lr = LinearRegression().fit(X, y)
            self.As[k], self.bs[k] = lr.coef_, lr.intercept_
            resid = y - lr.predict(X)
            Sigmas.append(np.cov(resid, rowvar=False).reshape(D, D) + 1e-8 * np.eye(D))
        self.Sigmas = np.array(Sigmas)

    def m_step(self, stats, counts):
        """Weighted least squares from the GPU-reduced Gram blocks.

        ``stats[k]`` is the Gram matrix of [phi_t ; x_t] with phi_t = [x_{t-1..t-L}, 1] weighted by
        gamma_t(k); ``counts[k]`` = sum_t gamma_t(k)."""
        K, D, L = self.K, self.D, self.lags
        P = D * L + 1
        J0 = np.diag(np.concatenate([self.l2_penalty_A * np.ones(D * L), [self.l2_penalty_b]]))
        for k in range(K):
            Sxx, Sxy, Syy = stats[k][:P, :P], stats[k][:P, P:], stats[k][P:, P:]
            W = np.linalg.solve(Sxx + J0, Sxy).T
            self.As[k], self.bs[k] = W[:, :D * L], W[:, -1]
            WSxy = W @ Sxy
            sqerr = Syy - WSxy.T - WSxy + W @ Sxx @ W.T
            S = (sqerr + self.Psi0) / (self.nu0 + counts[k] + D + 1)
            self.Sigmas[k] = 0.5 * (S + S.T)
        unused = np.where(counts < 1)[0]
        used = np.where(counts > 1)[0]
        for k in unused:
            if len(used) == 0:
                break
            i = np.random.choice(used)
            self.As[k] = self.As[i] + 0.01 * np.random.randn(*self.As[i].shape)
            self.bs[k] = self.bs[i] + 0.01 * np.random.randn(*self.bs[i].shape)
            self.Sigmas[k] = self.Sigmas[i]

    def sample_x(self, z, xhist, input=None, tag=None, with_noise=True):
        D, L = self.D, self.lags
        if xhist.shape[0] < L:
            mu = np.zeros(D)
            S = np.eye(D)
        else:
            mu = self.bs[z].copy()
            for l in range(L):
                mu += self.As[z][:, l * D:(l + 1) * D] @ xhist[-l - 1]
            S = self.Sigmas[z]
        if not with_noise:
            return mu
        return mu + np.linalg.cholesky(S) @ np.random.randn(D)


def _random_rotation(n, theta=None):
    """ssm.util.random_rotation."""
    theta = 0.5 * np.pi * np.random.rand() if theta is None else theta
    if n == 1:
        return np.random.rand() * np.eye(1)
    rot = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    out = np.eye(n)
    out[:2, :2] = rot
    q = np.linalg.qr(np.random.randn(n, n))[0]
    return q.dot(out).dot(q.T)


class _PinnedPool:
    """Grow-only pinned staging buffer shared by every HMM of the process.

    Allocating ~100 MB of page-locked memory costs tens of milliseconds, more than the E-step it
    feeds; the buffer is therefore kept and reused, guarded by an event recorded after the last
    copy that read from it."""
    buf = None
    event = None

    @classmethod
    def acquire(cls, n_floats):
        import torch
        if cls.event is not None:
            cls.event.synchronize()
            cls.event = None
        if cls.buf is None or cls.buf.numel() < n_floats:
            cls.buf = None
            cls.buf = torch.empty(max(n_floats, 1 << 20), dtype=torch.float32).pin_memory()
        return cls.buf

    @classmethod
    def release(cls):
        import torch
        cls.event = torch.cuda.Event()
        cls.event.record()


_STAGE_CHUNK_ROWS = 1 << 18          # rows per host-gather / H2D piece (12 MB at D = 12)


def _gather_threads():
    """Host threads of one rank's row gather: up to 8, but the ranks of a node share its cores (8 ranks x 8 threads on
    a 32-core host made the staging of every rank slower than one rank alone)."""
    cores = os.cpu_count() or 1
    local = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')) or 1)
    return max(1, min(8, cores // max(1, local)))


def _gather_rows(datas, lens, D, out):
    """Concatenate per-trial (T_i, D) host arrays into ``out`` ((sum T_i, D) float32) with the library's
    multi-threaded row gather (bn_host_gather_rows); arrays that are not C-contiguous float32 / float64
    are converted first."""
    keep = [d if (d.dtype == np.float32 or d.dtype == np.float64) and d.flags.c_contiguous
            else np.ascontiguousarray(d, dtype=np.float32) for d in datas]
    n = len(keep)
    ptrs = np.fromiter((d.__array_interface__['data'][0] for d in keep), dtype=np.uint64, count=n)
    rows = np.asarray(lens, dtype=np.int64)
    f64 = np.fromiter((d.dtype == np.float64 for d in keep), dtype=np.int32, count=n)
    _lib.check(_lib.lib().bn_host_gather_rows(
        ptrs.ctypes.data, rows.ctypes.data, f64.ctypes.data, n, D, out.ctypes.data,
        _gather_threads()), 'bn_host_gather_rows')


class _Staged:
    """Trials concatenated on the device: x (total_T, D) fp32, offsets (n+1) int64."""

    def __init__(self, datas, D, device):
        import torch
        lens = [int(d.shape[0]) for d in datas]
        self.n = len(datas)
        self.lengths = lens
        self.total = int(sum(lens))
        self.max_T = max(lens) if lens else 0
        self.x = torch.empty((self.total, D), dtype=torch.float32, device=device)
        if self.total:
            # the host arrays are gathered piece by piece into pinned memory and each piece is sent
            # as soon as it is complete, so the copy engine overlaps the host-side gather
            pinned = _PinnedPool.acquire(self.total * D)[:self.total * D].view(self.total, D)
            host = pinned.numpy()
            i = o = 0
            while i < self.n:
                j, rows = i, 0
                while j < self.n and (rows == 0 or rows + lens[j] <= _STAGE_CHUNK_ROWS):
                    rows += lens[j]
                    j += 1
                if rows:
                    _gather_rows(datas[i:j], lens[i:j], D, host[o:o + rows])
                    self.x[o:o + rows].copy_(pinned[o:o + rows], non_blocking=True)
                o += rows
                i = j
            _PinnedPool.release()
        off = np.zeros(self.n + 1, np.int64)
        off[1:] = np.cumsum(lens)
        self.offsets_host = off
        self.offsets = torch.from_numpy(off).to(device)

    @classmethod
    def from_device(cls, x, lengths):
        """Stage trials that already live on the GPU (e.g. encoder latents, config C5): ``x`` is the
        (sum T_i, D) fp32 concatenation, ``lengths`` the per-trial T_i.  No host round trip."""
        import torch
        self = cls.__new__(cls)
        lens = [int(v) for v in lengths]
        self.n = len(lens)
        self.lengths = lens
        self.total = int(sum(lens))
        self.max_T = max(lens) if lens else 0
        if x.dim() != 2 or x.shape[0] != self.total or x.dtype != torch.float32 or not x.is_cuda:
            raise ValueError('expected a CUDA float32 tensor of shape (%d, D)' % self.total)
        self.x = x.contiguous()
        off = np.zeros(self.n + 1, np.int64)
        off[1:] = np.cumsum(lens)
        self.offsets_host = off
        self.offsets = torch.from_numpy(off).to(x.device)
        return self


class HMM:
    """ssm.HMM look-alike (see module docstring)."""

    def __init__(self, K, D, M=0, init_state_distn=None, transitions='standard', transition_kwargs=None,
                 observations='gaussian', observation_kwargs=None, device=None, **kwargs):
        if M != 0:
            raise NotImplementedError('input-driven HMMs have no B200 kernel path')
        self.K, self.D, self.M = K, D, M
        self.init_state_distn = InitialStateDistribution(K)
        transition_kwargs = transition_kwargs or {}
        if transitions in ('standard', 'stationary'):
            self.transitions = StationaryTransitions(K)
        elif transitions == 'sticky':
            self.transitions = StickyTransitions(K, **transition_kwargs)
        else:
            raise NotImplementedError('transitions="%s" has no B200 kernel path' % transitions)
        observation_kwargs = observation_kwargs or {}
        if observations in ('ar', 'autoregressive'):
            self.observations = AutoRegressiveObservations(K, D, **observation_kwargs)
        elif observations == 'gaussian':
            self.observations = AutoRegressiveObservations(K, D, lags=0)
        else:
            raise NotImplementedError('observations="%s" has no B200 kernel path' % observations)
        self.device = device
        self.data_parallel = False

    # -- pickling: parameters only (device caches are rebuilt lazily) --------------------------
    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop('_cache', None)
        st.pop('_blob_cache', None)
        return st

    # -- ssm-style accessors --------------------------------------------------------------------
    @property
    def params(self):
        o = self.observations
        return ((self.init_state_distn.log_pi0,), (self.transitions.log_Ps,), (o.As, o.bs, o.Sigmas))

    def permute(self, perm):
        """Reorder the discrete states (arhmm_grid_search.py:204)."""
        perm = np.asarray(perm)
        assert np.all(np.sort(perm) == np.arange(self.K))
        self.init_state_distn.permute(perm)
        self.transitions.permute(perm)
        self.observations.permute(perm)

    def initialize(self, datas, inputs=None, masks=None, tags=None):
        """ssm.HMM.initialize: observation initialisation without localisation; BehaveNet
        immediately overrides it with ``observations.initialize(..., localize=True)``."""
        self.observations.initialize(_as_list(datas), localize=False)

    def log_prior(self):
        return self.transitions.log_prior() + self.observations.log_prior()

    # -- device plumbing --------------------------------------------------------------------------
    def _device(self):
        import torch
        if self.device is not None:
            return torch.device(self.device)
        if not torch.cuda.is_available():
            raise _lib.NativeLibraryError('no CUDA device: the ARHMM has no CPU path')
        return torch.device('cuda', torch.cuda.current_device())

    def _stage(self, datas):
        datas = _as_list(datas)
        # identity of the arrays + a content fingerprint (first / middle / last row of up to 64 evenly spaced
        # trials, plus the very first and last rows): reusing a list whose arrays were modified IN PLACE
        # (normalised, re-scaled, ...) re-stages; the cache keeps the arrays alive so their ids cannot be
        # recycled while it is valid.  ``clear_cache()`` drops it explicitly.
        key = tuple((id(d), d.shape[0]) for d in datas)
        if datas:
            import zlib
            crc = 0
            step = max(1, len(datas) // 64)
            for d in list(datas[::step]) + [datas[-1]]:
                if d.size:
                    T = d.shape[0]
                    crc = zlib.crc32(np.ascontiguousarray(d[[0, T // 2, T - 1]]).tobytes(), crc)
            key += (crc,)
        cache = self.__dict__.setdefault('_cache', {})
        if cache.get('key') != key:
            for d in datas:
                if d.ndim != 2 or d.shape[1] != self.D:
                    raise ValueError('data must have shape (T, %d), got %s' % (self.D, d.shape))
            cache.clear()
            cache['key'] = key
            cache['refs'] = datas
            cache['staged'] = _Staged(datas, self.D, self._device())
        return cache['staged']

    def clear_cache(self):
        self.__dict__.pop('_cache', None)

    def stage_device(self, x, lengths):
        """Device-resident trials for the E-step (see ``_Staged.from_device``)."""
        if x.shape[1] != self.D:
            raise ValueError('data must have shape (T, %d), got %s' % (self.D, tuple(x.shape)))
        return _Staged.from_device(x, lengths)

    def expected_states_device(self, staged):
        """E-step over device-resident trials: (Ez (sum T, K), Ezz (n, K, K), logZ (n)) as CUDA tensors."""
        return self._run_estep(staged, True)

    def _blob(self, device):
        """Device copy of the packed parameters (fp64 Cholesky whitening + TF32 hi/lo split on the host),
        cached by parameter CONTENT: an E-step, a log-likelihood and a Viterbi pass between two M-steps share
        one pack + one host->device copy instead of repeating both per call."""
        import torch
        import zlib
        o = self.observations
        K, D, L = self.K, self.D, o.lags
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (
            self.init_state_distn.log_pi0 - _logsumexp(self.init_state_distn.log_pi0),
            self.transitions.log_Ps - _logsumexp(self.transitions.log_Ps, axis=1, keepdims=True),
            o.As, o.bs, o.Sigmas)]
        crc = 0
        for a in arrs:
            crc = zlib.crc32(a.tobytes(), crc)
        key = (K, D, L, crc, sum(a.size for a in arrs), str(device))
        cached = self.__dict__.get('_blob_cache')
        if cached is not None and cached[0] == key:
            return cached[1]
        lib = _lib.lib()
        nbytes = lib.bn_arhmm_params_bytes(K, D, L)
        if nbytes == 0:
            raise _lib.NativeLibraryError(lib.bn_last_error().decode())
        host = np.zeros(nbytes, np.uint8)
        _lib.check(lib.bn_arhmm_pack_params(K, D, L, *[a.ctypes.data for a in arrs], host.ctypes.data),
                   'bn_arhmm_pack_params')
        blob = torch.from_numpy(host).to(device)
        self.__dict__['_blob_cache'] = (key, blob)
        return blob

    def _run_estep(self, st, want_post, shard=None):
        """Run the E-step kernels over trials [shard) and return device tensors."""
        import torch
        K, D, L = self.K, self.D, self.observations.lags
        dev = st.x.device
        lib = _lib.lib()
        lo, hi = shard if shard is not None else (0, st.n)
        n = hi - lo
        t0, t1 = int(st.offsets_host[lo]), int(st.offsets_host[hi])
        total = t1 - t0
        Ez = torch.empty(total, K, dtype=torch.float32, device=dev) if want_post else None
        Ezz = torch.zeros(n, K, K, dtype=torch.float32, device=dev) if want_post else None
        logZ = torch.zeros(n, dtype=torch.float64, device=dev)
        if n == 0 or total == 0:
            return Ez, Ezz, logZ
        offs = (st.offsets[lo:hi + 1] - t0).contiguous()
        x = st.x[t0:t1]
        max_T = max(st.lengths[lo:hi])
        blob = self._blob(dev)
        ws = torch.empty(lib.bn_arhmm_workspace_bytes(K, D, L, total, n, 0), dtype=torch.uint8, device=dev)
        _lib.check(lib.bn_arhmm_estep(
            K, D, L, blob.data_ptr(), x.data_ptr(), offs.data_ptr(), n, total, max_T, ws.data_ptr(),
            _lib.ptr(Ez), _lib.ptr(Ezz), logZ.data_ptr(), _lib.stream_ptr()), 'bn_arhmm_estep')
        return Ez, Ezz, logZ

    def _shard(self, n):
        if self.data_parallel and parallel.enabled():
            return parallel.shard_range(n)
        return 0, n

    # -- the hot-path calls --------------------------------------------------------------------
    def expected_states(self, data, input=None, mask=None, tag=None):
        """(Ez (T,K), Ezz (1,K,K) summed over time, log normaliser) -- ssm.HMM.expected_states."""
        st = self._stage([np.asarray(data)])
        Ez, Ezz, logZ = self._run_estep(st, True)
        return (Ez.cpu().numpy().astype(np.float64), Ezz.cpu().numpy().astype(np.float64),
                float(logZ.cpu().numpy()[0]))

    def log_likelihood(self, datas, inputs=None, masks=None, tags=None):
        """Sum of log normalisers; accepts one (T, D) array or a list (arhmm_grid_search.py:173,196)."""
        st = self._stage(datas)
        lo, hi = self._shard(st.n)
        _, _, logZ = self._run_estep(st, False, (lo, hi))
        tot = logZ.sum().reshape(1)
        if self.data_parallel and parallel.enabled():
            parallel.all_reduce_sum(tot)
        return float(tot.item())

    def log_probability(self, datas, inputs=None, masks=None, tags=None):
        return self.log_prior() + self.log_likelihood(datas)

    def most_likely_states(self, data, input=None, mask=None, tag=None):
        """Viterbi path of one trial (ssm.HMM.most_likely_states; arhmm_grid_search.py:201)."""
        return self.most_likely_states_batch([np.asarray(data)])[0]

    def most_likely_states_batch(self, datas):
        """Viterbi paths of many trials in one launch (what export_states loops over, eval.py:167)."""
        import torch
        st = self._stage(datas)
        K, D, L = self.K, self.D, self.observations.lags
        dev = st.x.device
        lib = _lib.lib()
        z = torch.zeros(st.total, dtype=torch.int32, device=dev)
        if st.total:
            blob = self._blob(dev)
            ws = torch.empty(lib.bn_arhmm_workspace_bytes(K, D, L, st.total, st.n, 1), dtype=torch.uint8,
                             device=dev)
            _lib.check(lib.bn_arhmm_viterbi(
                K, D, L, blob.data_ptr(), st.x.data_ptr(), st.offsets.data_ptr(), st.n, st.total,
                st.max_T, ws.data_ptr(), z.data_ptr(), _lib.stream_ptr()), 'bn_arhmm_viterbi')
        zh = z.cpu().numpy().astype(np.int64)
        return [zh[st.offsets_host[i]:st.offsets_host[i + 1]] for i in range(st.n)]

    def e_step_stats(self, datas):
        """One sharded E-step + sufficient statistics, all-reduced over ranks.

        Returns dict(gamma0 (K), joints (K,K), ar_stats (K,Q,Q), counts (K), ll float)."""
        import torch
        st = self._stage(datas)
        K, D, L = self.K, self.D, self.observations.lags
        Q = D * L + 1 + D
        dev = st.x.device
        lib = _lib.lib()
        lo, hi = self._shard(st.n)
        Ez, Ezz, logZ = self._run_estep(st, True, (lo, hi))
        flat = torch.zeros(K + K * K + K * Q * Q + K + 1, dtype=torch.float64, device=dev)
        n = hi - lo
        t0, t1 = int(st.offsets_host[lo]), int(st.offsets_host[hi])
        if n > 0 and t1 > t0:
            offs = (st.offsets[lo:hi + 1] - t0).contiguous()
            starts = offs[:-1][(offs[1:] - offs[:-1]) > 0]
            flat[:K] = Ez[starts].sum(0).double()
            flat[K:K + K * K] = Ezz.sum(0).double().reshape(-1)
            stats = flat[K + K * K:K + K * K + K * Q * Q]
            counts = flat[K + K * K + K * Q * Q:K + K * K + K * Q * Q + K]
            _lib.check(lib.bn_arhmm_ar_stats(
                K, D, L, st.x[t0:t1].data_ptr(), offs.data_ptr(), n, t1 - t0, Ez.data_ptr(),
                stats.data_ptr(), counts.data_ptr(), _lib.stream_ptr()), 'bn_arhmm_ar_stats')
            flat[-1] = logZ.sum()
        if self.data_parallel and parallel.enabled():
            parallel.all_reduce_sum(flat)       # the one collective of an EM iteration
        h = flat.cpu().numpy()
        o = K + K * K
        ar = h[o:o + K * Q * Q].reshape(K, Q, Q)
        ar = np.triu(ar) + np.transpose(np.triu(ar, 1), (0, 2, 1))
        return {'gamma0': h[:K], 'joints': h[K:o].reshape(K, K), 'ar_stats': ar,
                'counts': h[o + K * Q * Q:o + K * Q * Q + K], 'll': float(h[-1])}

    def fit(self, datas, inputs=None, masks=None, tags=None, method='em', num_iters=100,
            initialize=True, tolerance=0, verbose=0, **kwargs):
        """EM (ssm.HMM.fit(method='em')): returns the list of log probabilities, first entry before
        any update (arhmm_grid_search.py:170 calls it with num_iters=1, initialize=False)."""
        if method != 'em':
            raise NotImplementedError('only method="em" is supported')
        datas = _as_list(datas)
        if initialize:
            self.initialize(datas)
        lls = [self.log_probability(datas)]
        for _ in range(num_iters):
            s = self.e_step_stats(datas)
            self.init_state_distn.m_step(s['gamma0'])
            self.transitions.m_step(s['joints'])
            self.observations.m_step(s['ar_stats'], s['counts'])
            lls.append(self.log_prior() + s['ll'])
            if tolerance > 0 and len(lls) > 2 and abs(lls[-1] - lls[-2]) < tolerance:
                break
        return lls

    def sample(self, T, prefix=None, input=None, tag=None, with_noise=True):
        """(z, x) sample path (ssm.HMM.sample; plotting/arhmm_utils.py:235)."""
        K, D = self.K, self.D
        P = self.transitions.transition_matrix
        if prefix is None:
            z = np.zeros(T, dtype=int)
            x = np.zeros((T, D))
            z[0] = np.random.choice(K, p=self.init_state_distn.initial_state_distn)
            x[0] = self.observations.sample_x(z[0], x[:0], with_noise=with_noise)
            start = 1
        else:
            zpre, xpre = prefix
            start = len(zpre)
            z = np.concatenate([zpre, np.zeros(T, dtype=int)])
            x = np.concatenate([xpre, np.zeros((T, D))])
            T = T + start
        for t in range(start, T):
            z[t] = np.random.choice(K, p=P[z[t - 1]])
            x[t] = self.observations.sample_x(z[t], x[:t], with_noise=with_noise)
        if prefix is None:
            return z, x
        return z[start:], x[start:]


def _as_list(datas):
    if isinstance(datas, np.ndarray) and datas.ndim == 2:
        return [datas]
    if isinstance(datas, (list, tuple)):
        return [np.asarray(d) for d in datas]
    raise TypeError('datas must be a (T, D) array or a list of them')
