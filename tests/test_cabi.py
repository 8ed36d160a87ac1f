"""CPU tests of the C-ABI boundary: the shared library loads, exports every symbol that
include/behavenet_b200.h declares, the ctypes struct mirrors the C struct, and the host-only
entry points (no GPU needed) behave: parameter packing of the ARHMM and argument validation."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from behavenet_b200 import _lib
from oracle import arhmm_oracle as ao

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'behavenet_b200.h')


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bn_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), 'missing export: ' + name
        assert name in _lib.SIGNATURES, 'no ctypes signature for ' + name
    assert sorted(_lib.SIGNATURES) == names
    assert lib.bn_abi_version() == 3


def test_desc_struct_matches_header():
    text = open(HEADER).read()
    body = text[text.index('typedef struct bn_cae_desc {'):text.index('} bn_cae_desc;')]
    scalars = len(re.findall(r'int32_t\s+[a-z0-9_, ]+;', body))
    assert scalars > 0
    n_arrays = len(re.findall(r'\[BN_MAX_LAYERS\]', body))
    n_scalars = 6 + 3
    assert C.sizeof(_lib.CaeDesc) == 4 * (n_scalars + n_arrays * _lib.BN_MAX_LAYERS)
    assert [f[0] for f in _lib.CaeDesc._fields_][:6] == [
        'n_layers', 'in_c', 'in_h', 'in_w', 'n_latents', 'n_heads']


def test_tensor_core_mode_switch():
    lib = _lib.lib()
    old = lib.bn_get_tensor_core_mode()
    lib.bn_set_tensor_core_mode(0)
    assert lib.bn_get_tensor_core_mode() == 0
    lib.bn_set_tensor_core_mode(1)
    assert lib.bn_get_tensor_core_mode() == 1
    lib.bn_set_tensor_core_mode(old)
    assert lib.bn_launch_count() >= 0


def _blob(p):
    lib = _lib.lib()
    n = lib.bn_arhmm_params_bytes(p.K, p.D, p.lags)
    assert n > 0
    buf = np.zeros(n, np.uint8)
    arrs = [np.ascontiguousarray(a, np.float64) for a in (p.log_pi0, p.log_Ps, p.As, p.bs, p.Sigmas)]
    _lib.check(lib.bn_arhmm_pack_params(p.K, p.D, p.lags, *[a.ctypes.data for a in arrs],
                                        buf.ctypes.data), 'pack')
    return buf


@pytest.mark.parametrize('K,D,lags', [(16, 12, 2), (3, 5, 1), (4, 3, 0)])
def test_arhmm_param_blob_whitening(K, D, lags):
    """The host-side packer folds chol(Sigma)^-1 [I, -A, -b] per state; evaluating the folded
    form in numpy must reproduce the oracle's emission log-likelihoods."""
    p = ao.synth_params(K, D, lags, seed=3)
    buf = _blob(p)
    hdr = np.frombuffer(buf[:32].tobytes(), np.int32)
    assert list(hdr[:3]) == [K, D, lags]
    DP, J, KP = int(hdr[3]), int(hdr[4]), int(hdr[5])
    assert DP % 4 == 0 and DP >= D and J == D * (lags + 1) + 1 and KP >= K
    offs = np.frombuffer(buf[32:32 + 9 * 8].tobytes(), np.int64)
    off_Wd, off_cd = int(offs[6]), int(offs[7])
    W = np.frombuffer(buf[off_Wd:off_Wd + 8 * K * J * DP].tobytes(), np.float64).reshape(K, J, DP)
    c = np.frombuffer(buf[off_cd:off_cd + 8 * (K + 1)].tobytes(), np.float64)
    x = ao.sample(p, 40, np.random.RandomState(0))[1]
    ll_ref = ao.ar_log_likelihoods(x, p.As, p.bs, p.Sigmas, p.lags)
    for t in range(lags, 40):
        psi = np.concatenate([x[t]] + [x[t - l - 1] for l in range(lags)] + [[1.0]])
        for k in range(K):
            y = psi @ W[k][:, :D]
            assert abs(c[k] - 0.5 * (y ** 2).sum() - ll_ref[t, k]) < 1e-8
    # float copies are the rounded doubles, transition rows are probabilities
    off_Pf = int(offs[1])
    P = np.frombuffer(buf[off_Pf:off_Pf + 4 * KP * KP].tobytes(), np.float32).reshape(KP, KP)
    np.testing.assert_allclose(P[:K, :K], np.exp(p.log_Ps), rtol=1e-6)
    assert np.all(P[K:] == 0) and np.all(P[:, K:] == 0)


def test_argument_validation_reports_errors():
    lib = _lib.lib()
    assert lib.bn_arhmm_params_bytes(64, 4, 1) == 0          # K > 32 unsupported
    assert b'K=64' in lib.bn_last_error()
    p = ao.synth_params(3, 2, 1, seed=0)
    bad = p.Sigmas.copy()
    bad[1] = -np.eye(2)
    buf = np.zeros(lib.bn_arhmm_params_bytes(3, 2, 1), np.uint8)
    arrs = [np.ascontiguousarray(a, np.float64) for a in (p.log_pi0, p.log_Ps, p.As, p.bs, bad)]
    rc = lib.bn_arhmm_pack_params(3, 2, 1, *[a.ctypes.data for a in arrs], buf.ctypes.data)
    assert rc != 0 and b'positive definite' in lib.bn_last_error()
    with pytest.raises(_lib.NativeLibraryError):
        _lib.check(rc, 'pack')
    d = _lib.CaeDesc()
    d.n_layers = 0
    out = C.c_void_p()
    assert lib.bn_cae_plan_create(C.byref(d), C.byref(out)) != 0
    assert b'n_layers' in lib.bn_last_error()
    # linear autoencoder entry points: validated before any launch
    assert lib.bn_linae_workspace_bytes(256, 16384, 12) == 4 * (2 * 256 * 12 + 256 * 16384)
    assert lib.bn_linae_workspace_bytes(0, 16384, 12) == 0
    fake = C.c_void_p(1 << 20)      # never dereferenced: the calls below fail on their arguments
    assert lib.bn_linae_forward(4, 100, 65, fake, fake, fake, fake, fake, None, None) != 0
    assert b'L=65' in lib.bn_last_error()
    assert lib.bn_linae_forward(4, 100, 8, None, fake, fake, fake, fake, None, None) != 0
    assert b'null argument' in lib.bn_last_error()
    assert lib.bn_linae_loss(4, 100, 8, fake, None, fake, fake, fake, 200, 3, 6, fake, fake, None, None, None, None) != 0
    assert b'outside the batch' in lib.bn_last_error()
    assert lib.bn_linae_decode(0, 100, 8, fake, fake, fake, fake, None) == 0        # nothing to do
    # two-phase encoder backward: the phase selector is checked with the other arguments
    assert lib.bn_cae_encode_bwd_phase(fake, 4, fake, fake, None, fake, fake, fake, fake, None, 7) != 0
    assert b'phase 7' in lib.bn_last_error()


def test_host_gather_rows_matches_concatenate():
    """bn_host_gather_rows (host-only): ragged fp32 / fp64 trials, empty trials, thread counts."""
    import ctypes as C
    lib = _lib.lib()
    rng = np.random.RandomState(0)
    for D, lens, threads in [(12, [5, 0, 1000, 3, 77], 4), (3, [1], 1), (7, [40000, 1, 0, 65000, 12], 8), (5, [], 2)]:
        arrs = [rng.randn(T, D).astype(np.float64 if i % 2 else np.float32) for i, T in enumerate(lens)]
        n = len(arrs)
        ptrs = np.array([a.ctypes.data for a in arrs], dtype=np.uint64)
        rows = np.array(lens, dtype=np.int64)
        f64 = np.array([a.dtype == np.float64 for a in arrs], dtype=np.int32)
        out = np.full((int(sum(lens)), D), np.nan, np.float32)
        rc = lib.bn_host_gather_rows(ptrs.ctypes.data if n else None, rows.ctypes.data if n else None,
                                     f64.ctypes.data if n else None, n, D, out.ctypes.data if out.size else None,
                                     threads)
        assert rc == 0, lib.bn_last_error()
        ref = np.concatenate([a.astype(np.float32) for a in arrs], 0) if n else np.zeros((0, D), np.float32)
        assert np.array_equal(out, ref)
    assert lib.bn_host_gather_rows(None, None, None, 3, 4, None, 1) < 0
