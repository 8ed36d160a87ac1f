"""Layer-geometry calculator for the convolutional autoencoder hot path.

Host-side mirror of the part of the reference's architecture generator that the CAE hot path
consumes through ``hparams`` (reference: ``behavenet/models/ae_model_architecture_generator.py``
``calculate_output_dim`` 347-410, ``get_decoding_conv_block`` 271-344, ``get_handcrafted_dims``
482-592, ``load_handcrafted_arch`` 595-662, ``load_default_arch`` 707-720).  The random
architecture search / memory-footprint estimator is out of scope.

Only strided-conv architectures are accepted: max-pool layers, batch-norm and 'valid' padding
have no sm_100a kernel and raise ``NotImplementedError`` (there is no eager fallback).
"""

import copy
import json
import re

__all__ = [
    'load_default_arch', 'calculate_output_dim', 'get_decoding_conv_block',
    'get_handcrafted_dims', 'load_handcrafted_arch', 'load_handcrafted_arches']


def load_default_arch():
    """Default 5-layer strided CAE (reference ``load_default_arch`` 707-720)."""
    n_layers = 5
    return {
        'ae_network_type': 'strides_only',
        'ae_padding_type': 'same',
        'ae_batch_norm': 0,
        'ae_batch_norm_momentum': None,
        'symmetric_arch': 1,
        'ae_encoding_n_channels': [32 << i for i in range(n_layers)],
        'ae_encoding_kernel_size': [5] * n_layers,
        'ae_encoding_stride_size': [2, 2, 2, 2, 5],
        'ae_encoding_layer_type': ['conv'] * n_layers,
        'ae_decoding_last_FF_layer': 0}


def calculate_output_dim(input_dim, kernel, stride, padding_type, layer_type):
    """Output size and (before, after) zero padding of one spatial dim (reference 347-410).

    'same' follows the TensorFlow convention: out = ceil(in / stride), total padding
    ``max(0, (out-1)*stride + kernel - in)`` split with the smaller half first.
    """
    if layer_type == 'conv':
        if padding_type == 'same':
            out = -(-input_dim // stride)
            total = max(0, (out - 1) * stride + kernel - input_dim)
            return out, total // 2, total - total // 2
        if padding_type == 'valid':
            return (input_dim - kernel) // stride + 1, 0, 0
        raise NotImplementedError
    if layer_type == 'maxpool':
        if kernel != 2:
            raise NotImplementedError
        span = input_dim - kernel
        if padding_type == 'same':
            return -(-span // stride) + 1, 0, 0
        if padding_type == 'valid':
            return span // stride + 1, 0, 0
        raise NotImplementedError
    raise NotImplementedError


def get_decoding_conv_block(arch):
    """Mirror the encoder into a transposed-conv decoder (reference 271-344)."""
    enc_c = arch['ae_encoding_n_channels']
    n = len(enc_c)
    arch['ae_decoding_starting_dim'] = [
        enc_c[-1], arch['ae_encoding_y_dim'][-1], arch['ae_encoding_x_dim'][-1]]
    for key in ['x_dim', 'y_dim', 'x_padding', 'y_padding', 'n_channels', 'kernel_size',
                'stride_size', 'layer_type']:
        arch['ae_decoding_' + key] = []
    for src in range(n - 1, -1, -1):
        first = src == 0
        arch['ae_decoding_n_channels'].append(arch['ae_input_dim'][0] if first else enc_c[src - 1])
        arch['ae_decoding_kernel_size'].append(arch['ae_encoding_kernel_size'][src])
        arch['ae_decoding_stride_size'].append(arch['ae_encoding_stride_size'][src])
        arch['ae_decoding_x_padding'].append(arch['ae_encoding_x_padding'][src])
        arch['ae_decoding_y_padding'].append(arch['ae_encoding_y_padding'][src])
        arch['ae_decoding_y_dim'].append(
            arch['ae_input_dim'][1] if first else arch['ae_encoding_y_dim'][src - 1])
        arch['ae_decoding_x_dim'].append(
            arch['ae_input_dim'][2] if first else arch['ae_encoding_x_dim'][src - 1])
        kind = arch['ae_encoding_layer_type'][src]
        arch['ae_decoding_layer_type'].append({'maxpool': 'unpool', 'conv': 'convtranspose'}[kind])
    if arch['ae_decoding_last_FF_layer']:
        arch['ae_decoding_n_channels'][-1] = 16
    return arch


def get_handcrafted_dims(arch, symmetric=True):
    """Fill per-layer dims/padding lists of a handcrafted arch (reference 482-592)."""
    arch['model_type'] = 'conv'
    for key in ['x_dim', 'y_dim', 'x_padding', 'y_padding']:
        arch['ae_encoding_' + key] = []
    in_y, in_x = arch['ae_input_dim'][1], arch['ae_input_dim'][2]
    for i in range(len(arch['ae_encoding_n_channels'])):
        k = arch['ae_encoding_kernel_size'][i]
        s = arch['ae_encoding_stride_size'][i]
        kind = arch['ae_encoding_layer_type'][i]
        out_x, x0, x1 = calculate_output_dim(in_x, k, s, arch['ae_padding_type'], kind)
        out_y, y0, y1 = calculate_output_dim(in_y, k, s, arch['ae_padding_type'], kind)
        arch['ae_encoding_x_dim'].append(out_x)
        arch['ae_encoding_y_dim'].append(out_y)
        arch['ae_encoding_x_padding'].append((x0, x1))
        arch['ae_encoding_y_padding'].append((y0, y1))
        in_y, in_x = out_y, out_x
    if symmetric:
        return get_decoding_conv_block(arch)
    # non-symmetric decoders: reproduce the reference's 'same' rule, including its use of the
    # x total for the trailing y pad (reference 548-556)
    if arch.get('ae_network_type') == 'max_pooling' or 'unpool' in arch['ae_decoding_layer_type']:
        raise NotImplementedError
    for key in ['x_dim', 'y_dim', 'x_padding', 'y_padding']:
        arch['ae_decoding_' + key] = []
    in_y, in_x = arch['ae_decoding_starting_dim'][1], arch['ae_decoding_starting_dim'][2]
    for i in range(len(arch['ae_decoding_n_channels'])):
        k = arch['ae_decoding_kernel_size'][i]
        s = arch['ae_decoding_stride_size'][i]
        if arch['ae_padding_type'] == 'valid':
            continue
        if arch['ae_padding_type'] != 'same':
            raise NotImplementedError
        out_x = in_x * s - s + 1
        tot_x = max(0, (in_x - 1) * s + k - out_x)
        out_y = in_y * s - s + 1
        tot_y = max(0, (in_y - 1) * s + k - out_y)
        arch['ae_decoding_x_dim'].append(out_x)
        arch['ae_decoding_y_dim'].append(out_y)
        arch['ae_decoding_x_padding'].append((tot_x // 2, tot_x - tot_x // 2))
        arch['ae_decoding_y_padding'].append((tot_y // 2, tot_x - tot_y // 2))
        in_y, in_x = out_y, out_x
    return arch


def _load_commented_json(path):
    """json with '#' / '//' line comments (the reference uses commentjson for its configs)."""
    with open(path, 'r') as f:
        text = f.read()
    text = re.sub(r'^\s*(#|//).*$', '', text, flags=re.M)
    text = re.sub(r'\s+#[^"\n]*$', '', text, flags=re.M)
    return json.loads(text)


def load_handcrafted_arch(
        input_dim, n_ae_latents, ae_arch_json, batch_size=None, check_memory=False,
        mem_limit_gb=10):
    """Arch dict for an input size (reference ``load_handcrafted_arch`` 595-662).

    ``check_memory`` is accepted for signature compatibility; the footprint estimator is not part
    of the hot path (a 180 GB part fits every handcrafted arch at the reference's batch sizes).
    """
    if ae_arch_json is None:
        arch = load_default_arch()
    else:
        try:
            arch = _load_commented_json(ae_arch_json)
        except FileNotFoundError:
            print('Warning! could not find ae arch defined in %s; using default architecture'
                  % ae_arch_json)
            arch = load_default_arch()
    arch['ae_batch_norm'] = arch['ae_batch_norm'] == 1
    input_dim = list(input_dim)
    arch['n_input_channels'], arch['y_pixels'], arch['x_pixels'] = input_dim
    arch['ae_input_dim'] = input_dim
    arch['n_ae_latents'] = n_ae_latents
    return get_handcrafted_dims(arch, symmetric=arch['symmetric_arch'] == 1)


def load_handcrafted_arches(input_dim, n_ae_latents, ae_arch_json, batch_size=None,
                            check_memory=False, mem_limit_gb=10):
    """List of arch dicts, one per latent count (reference 665-704)."""
    if isinstance(n_ae_latents, int):
        n_ae_latents = [n_ae_latents]
    elif isinstance(n_ae_latents, str):
        n_ae_latents = [int(v) for v in n_ae_latents.strip('[]').split(',')]
    return [load_handcrafted_arch(copy.copy(input_dim), n, ae_arch_json, batch_size,
                                  check_memory, mem_limit_gb) for n in n_ae_latents]
