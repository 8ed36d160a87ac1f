"""ctypes binding of libbehavenet_b200.so (the C ABI declared in include/behavenet_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception
is raised.  The product never computes on the CPU or through eager PyTorch.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbehavenet_b200.so')
BN_MAX_LAYERS = 8

_lib = None


class NativeLibraryError(RuntimeError):
    pass


_ARR = C.c_int32 * BN_MAX_LAYERS


class CaeDesc(C.Structure):
    """Mirror of ``bn_cae_desc`` (include/behavenet_b200.h)."""
    _fields_ = (
        [(k, C.c_int32) for k in ('n_layers', 'in_c', 'in_h', 'in_w', 'n_latents', 'n_heads')]
        + [(k, _ARR) for k in ('enc_c', 'enc_k', 'enc_s', 'enc_h', 'enc_w',
                               'enc_pt', 'enc_pb', 'enc_pl', 'enc_pr')]
        + [(k, C.c_int32) for k in ('dec_c0', 'dec_h0', 'dec_w0')]
        + [(k, _ARR) for k in ('dec_c', 'dec_k', 'dec_s', 'dec_h', 'dec_w',
                               'dec_pt', 'dec_pb', 'dec_pl', 'dec_pr')])


_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/behavenet_b200.h declares
SIGNATURES = {
    'bn_abi_version': (_i, []),
    'bn_last_error': (C.c_char_p, []),
    'bn_launch_count': (_i64, []),
    'bn_set_tensor_core_mode': (_i, [_i]),
    'bn_get_tensor_core_mode': (_i, []),
    'bn_cae_plan_create': (_i, [C.POINTER(CaeDesc), C.POINTER(_vp)]),
    'bn_cae_plan_destroy': (None, [_vp]),
    'bn_cae_packed_bytes': (_sz, [_vp]),
    'bn_cae_workspace_bytes': (_sz, [_vp, _i]),
    'bn_cae_pack_params': (_i, [_vp, _vp, _vp, _vp]),
    'bn_cae_encode': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bn_cae_encode_u8': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bn_cae_decode': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    'bn_cae_decode_bwd': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bn_cae_encode_bwd': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bn_cae_encode_bwd_phase': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    'bn_cae_layer_op': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bn_linae_workspace_bytes': (_sz, [_i, _i, _i]),
    'bn_linae_forward': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bn_linae_decode': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'bn_linae_loss': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bn_psvae_latent_workspace_bytes': (_sz, [_i, _i]),
    'bn_psvae_latent': (_i, [_i, _i, _i] + [_vp] * 9 + [_f, _f, _f, _f] + [_vp] * 11),
    'bn_psvae_latent_bwd': (_i, [_i, _i, _i] + [_vp] * 11),
    'bn_arhmm_params_bytes': (_sz, [_i, _i, _i]),
    'bn_arhmm_pack_params': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'bn_arhmm_workspace_bytes': (_sz, [_i, _i, _i, _i64, _i, _i]),
    'bn_arhmm_estep': (_i, [_i, _i, _i, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    'bn_arhmm_viterbi': (_i, [_i, _i, _i, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp]),
    'bn_host_gather_rows': (_i, [_vp, _vp, _vp, _i, _i, _vp, _i]),
    'bn_arhmm_ar_stats': (_i, [_i, _i, _i, _vp, _vp, _i, _i64, _vp, _vp, _vp, _vp]),
}


def lib():
    """Load (once) and return the shared library; raise NativeLibraryError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                '%s not found: build it with `python -m behavenet_b200.csrc.build` '
                '(or __graft_entry__.build()); there is no CPU / eager fallback' % LIB_PATH)
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:
            raise NativeLibraryError('cannot load %s: %s' % (LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.bn_abi_version() != 3:
            raise NativeLibraryError('ABI version mismatch')
        _lib = handle
    return _lib


def check(status, what=''):
    if status != 0:
        msg = lib().bn_last_error()
        raise NativeLibraryError('%s failed: %s' % (what, msg.decode() if msg else status))


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().bn_launch_count())
