"""Host-side driver of the CAE plan: geometry -> bn_cae_desc, parameter tables, packed-weight cache,
workspaces, and the autograd bridges used by ConvAEEncoder / ConvAEDecoder.

All compute happens in libbehavenet_b200.so; this file only moves pointers.  torch is used for
device memory, streams and autograd bookkeeping.
"""

import ctypes as C

import torch

from behavenet_b200 import _lib

_PLANS = {}     # desc key -> plan handle (never destroyed: a handful per process, a few KB each)


def _unsupported(hparams):
    """Variants of the reference architecture that have no sm_100a kernel raise here -- there is
    no eager fallback."""
    if hparams.get('model_type', 'conv') != 'conv':
        raise NotImplementedError('only model_type="conv" has a B200 kernel path')
    if hparams.get('ae_batch_norm', False):
        raise NotImplementedError('ae_batch_norm is not supported by the B200 kernels')
    if hparams.get('ae_decoding_last_FF_layer', False):
        raise NotImplementedError('ae_decoding_last_FF_layer is not supported by the B200 kernels')
    if hparams.get('ae_padding_type', 'same') not in ('same', 'valid'):
        raise ValueError('"%s" is not a valid padding type' % hparams['ae_padding_type'])
    if any(t != 'conv' for t in hparams['ae_encoding_layer_type']):
        raise NotImplementedError('max-pool encoders are not supported by the B200 kernels')
    if any(t != 'convtranspose' for t in hparams['ae_decoding_layer_type']):
        raise NotImplementedError('unpool decoders are not supported by the B200 kernels')


def make_desc(hparams, role=None):
    """bn_cae_desc from the reference's hparams lists (aes.py:25-36, 229-242).

    role='encoder': a conditional encoder sees n_labels/2 one-hot label images next to the frame's
    channels (aes.py:129-137).  role='decoder': the decoder's FF layer reads hparams['hidden_layer_size']
    values, which exceeds n_ae_latents when labels join the latents (ConditionalAE, aes.py:803)."""
    _unsupported(hparams)
    d = _lib.CaeDesc()
    n = len(hparams['ae_encoding_n_channels'])
    if n != len(hparams['ae_decoding_n_channels']) or n > _lib.BN_MAX_LAYERS:
        raise NotImplementedError('encoder/decoder depth %d unsupported' % n)
    d.n_layers = n
    d.in_c, d.in_h, d.in_w = [int(v) for v in hparams['ae_input_dim']]
    d.n_latents = int(hparams['n_ae_latents'])
    if role == 'encoder' and hparams.get('model_class') == 'cond-ae' and hparams.get('conditional_encoder', False):
        d.in_c += int(hparams['n_labels']) // 2
    if role == 'decoder':
        d.n_latents = int(hparams.get('hidden_layer_size', hparams['n_ae_latents']))
    d.n_heads = 2 if hparams.get('variational', False) else 1
    for i in range(n):
        d.enc_c[i] = int(hparams['ae_encoding_n_channels'][i])
        d.enc_k[i] = int(hparams['ae_encoding_kernel_size'][i])
        d.enc_s[i] = int(hparams['ae_encoding_stride_size'][i])
        d.enc_h[i] = int(hparams['ae_encoding_y_dim'][i])
        d.enc_w[i] = int(hparams['ae_encoding_x_dim'][i])
        d.enc_pt[i], d.enc_pb[i] = [int(v) for v in hparams['ae_encoding_y_padding'][i]]
        d.enc_pl[i], d.enc_pr[i] = [int(v) for v in hparams['ae_encoding_x_padding'][i]]
        d.dec_c[i] = int(hparams['ae_decoding_n_channels'][i])
        d.dec_k[i] = int(hparams['ae_decoding_kernel_size'][i])
        d.dec_s[i] = int(hparams['ae_decoding_stride_size'][i])
        d.dec_h[i] = int(hparams['ae_decoding_y_dim'][i])
        d.dec_w[i] = int(hparams['ae_decoding_x_dim'][i])
        d.dec_pt[i], d.dec_pb[i] = [int(v) for v in hparams['ae_decoding_y_padding'][i]]
        d.dec_pl[i], d.dec_pr[i] = [int(v) for v in hparams['ae_decoding_x_padding'][i]]
    d.dec_c0, d.dec_h0, d.dec_w0 = [int(v) for v in hparams['ae_decoding_starting_dim']]
    if hparams.get('ae_padding_type', 'same') == 'valid':
        # aes.py:382-405: ConvTranspose2d(padding=(y0, x0), output_padding=target size - full size).  In the
        # plan's terms (full transposed conv, then crop) the rows / columns that output_padding appends are a
        # NEGATIVE crop at the bottom / right: they lie outside every tap's reach and receive the bias only.
        h, w = d.dec_h0, d.dec_w0
        for i in range(n):
            if d.dec_pt[i] or d.dec_pl[i]:
                raise ValueError("decoder layer %d: 'valid' padding comes with zero pads" % i)
            op_y = d.dec_h[i] - ((h - 1) * d.dec_s[i] + d.dec_k[i])
            op_x = d.dec_w[i] - ((w - 1) * d.dec_s[i] + d.dec_k[i])
            if not (0 <= op_y < d.dec_s[i] and 0 <= op_x < d.dec_s[i]):
                raise ValueError('decoder layer %d: output_padding (%d, %d) outside [0, stride)' % (i, op_y, op_x))
            d.dec_pb[i], d.dec_pr[i] = -op_y, -op_x
            h, w = d.dec_h[i], d.dec_w[i]
    return d


def output_padding(hparams, i):
    """(y, x) ``output_padding`` of decoder layer i under 'valid' padding (aes.py:382-403)."""
    h0, w0 = (hparams['ae_decoding_starting_dim'][1:] if i == 0 else
              (hparams['ae_decoding_y_dim'][i - 1], hparams['ae_decoding_x_dim'][i - 1]))
    s, k = hparams['ae_decoding_stride_size'][i], hparams['ae_decoding_kernel_size'][i]
    return (hparams['ae_decoding_y_dim'][i] - ((h0 - 1) * s + k), hparams['ae_decoding_x_dim'][i] - ((w0 - 1) * s + k))


def desc_key(d):
    return bytes(d)


def get_plan(d, device):
    key = (desc_key(d), str(device))
    if key not in _PLANS:
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().bn_cae_plan_create(C.byref(d), C.byref(handle)), 'bn_cae_plan_create')
        _PLANS[key] = handle.value
    return _PLANS[key]


class Runtime:
    """Per-module cache of device buffers (packed weights, workspaces).  Holds no C handles, so
    modules stay deep-copyable / picklable (training.py:393-396 deep-copies the best model)."""

    def __init__(self):
        self.clear()

    def clear(self):
        self.packed = None
        self.packed_key = None
        self.ws = {}
        self.bufs = {}

    def __deepcopy__(self, memo):
        return Runtime()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.clear()


class CaeDriver:
    """Binds a geometry (desc) + a list of parameters (ordered like the C parameter table, None
    for absent entries) to the C calls."""

    def __init__(self, hparams, role=None):
        self.desc = make_desc(hparams, role)
        self.n_layers = self.desc.n_layers
        self.n_params = 4 * self.n_layers + 6
        self.L = self.desc.n_latents
        self.img = (self.desc.in_c, self.desc.in_h, self.desc.in_w)

    def __deepcopy__(self, memo):
        new = CaeDriver.__new__(CaeDriver)
        new.__dict__.update(self.__dict__)
        d = _lib.CaeDesc()
        C.memmove(C.byref(d), C.byref(self.desc), C.sizeof(d))
        new.desc = d
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st['desc'] = bytes(self.desc)
        return st

    def __setstate__(self, st):
        raw = st.pop('desc')
        self.__dict__.update(st)
        self.desc = _lib.CaeDesc.from_buffer_copy(raw)

    # -- helpers -----------------------------------------------------------------------------
    @staticmethod
    def _check_input(x, what, shape_tail=None, allow_uint8=False):
        if not x.is_cuda:
            raise RuntimeError('%s must live on a CUDA device (B200 kernels only; no CPU path)' % what)
        if x.dtype != torch.float32 and not (allow_uint8 and x.dtype == torch.uint8):
            raise TypeError('%s must be float32%s, got %s' % (what, ' or uint8' if allow_uint8 else '', x.dtype))
        if shape_tail is not None and tuple(x.shape[1:]) != tuple(shape_tail):
            raise ValueError('%s has shape %s, expected (n, %s)' % (
                what, tuple(x.shape), ', '.join(str(s) for s in shape_tail)))
        return x.contiguous()

    def plan(self, device):
        return get_plan(self.desc, device)

    def table(self, tensors):
        arr = (C.c_void_p * self.n_params)()
        for i, t in enumerate(tensors):
            arr[i] = None if t is None else t.data_ptr()
        return arr

    def packed(self, rt, params, device):
        """Packed-weight cache, rebuilt when any parameter changed (optimizer step,
        load_state_dict, .to())."""
        key = tuple((p.data_ptr(), p._version) for p in params if p is not None)
        if rt.packed is None or rt.packed_key != key or rt.packed.device != device:
            lib = _lib.lib()
            plan = self.plan(device)
            nbytes = lib.bn_cae_packed_bytes(plan)
            if rt.packed is None or rt.packed.numel() != nbytes or rt.packed.device != device:
                rt.packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
            _lib.check(lib.bn_cae_pack_params(plan, self.table(params), rt.packed.data_ptr(),
                                              _lib.stream_ptr()), 'bn_cae_pack_params')
            rt.packed_key = key
        return rt.packed

    def workspace(self, rt, n, device, fresh=False):
        key = (n, str(device))
        if fresh or key not in rt.ws:
            nbytes = _lib.lib().bn_cae_workspace_bytes(self.plan(device), n)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            if fresh:
                return ws
            rt.ws = {key: ws}      # keep only the latest batch size
        return rt.ws[key]

    # -- raw calls ---------------------------------------------------------------------------
    def encode(self, x, params, packed, ws, want_logvar):
        n = x.shape[0]
        mu = torch.empty(n, self.L, dtype=torch.float32, device=x.device)
        logvar = torch.empty(n, self.L, dtype=torch.float32, device=x.device) if want_logvar else None
        # raw uint8 video is scaled by 1/255 inside the first layer's loader (encode-only path)
        name = 'bn_cae_encode_u8' if x.dtype == torch.uint8 else 'bn_cae_encode'
        _lib.check(getattr(_lib.lib(), name)(
            self.plan(x.device), n, x.data_ptr(), self.table(params), packed.data_ptr(),
            ws.data_ptr(), mu.data_ptr(), _lib.ptr(logvar), _lib.stream_ptr()), name)
        return mu, logvar

    def decode(self, z, params, packed, ws, want_xhat=True, target=None, mask=None, chunk_size=0,
               frame_offset=0, n_total=0, grad_coef=0.0, sse=None):
        n = z.shape[0]
        xhat = torch.empty((n,) + self.img, dtype=torch.float32, device=z.device) if want_xhat else None
        _lib.check(_lib.lib().bn_cae_decode(
            self.plan(z.device), n, z.data_ptr(), self.table(params), packed.data_ptr(),
            ws.data_ptr(), _lib.ptr(xhat), _lib.ptr(target), _lib.ptr(mask), int(chunk_size),
            int(frame_offset), int(n_total), float(grad_coef), _lib.ptr(sse), _lib.stream_ptr()),
            'bn_cae_decode')
        return xhat

    def decode_bwd(self, n, dxhat, params, packed, ws, grads, device, want_dz=True):
        dz = torch.empty(n, self.L, dtype=torch.float32, device=device) if want_dz else None
        _lib.check(_lib.lib().bn_cae_decode_bwd(
            self.plan(device), n, _lib.ptr(dxhat), self.table(params), packed.data_ptr(),
            ws.data_ptr(), self.table(grads), _lib.ptr(dz), _lib.stream_ptr()), 'bn_cae_decode_bwd')
        return dz

    def encode_bwd(self, x, dmu, dlogvar, params, packed, ws, grads, phase=0):
        """phase 0: the whole pass; 1: heads + top conv layer (their gradients are final afterwards); 2: the
        layers below -- the split lets a data-parallel caller all-reduce the top bucket under phase 2."""
        _lib.check(_lib.lib().bn_cae_encode_bwd_phase(
            self.plan(x.device), x.shape[0], x.data_ptr(), _lib.ptr(dmu), _lib.ptr(dlogvar),
            self.table(params), packed.data_ptr(), ws.data_ptr(), self.table(grads),
            _lib.stream_ptr(), int(phase)), 'bn_cae_encode_bwd')


class EncodeFn(torch.autograd.Function):
    """Differentiable ConvAEEncoder.forward: (x, *encoder params) -> mu [, logvar]."""

    @staticmethod
    def forward(ctx, module, grad_on, x, *params):
        # grad_on = torch.is_grad_enabled() of the caller: inside Function.forward grad mode is always
        # off, and needs_input_grad ignores torch.no_grad()
        drv, rt = module._driver, module._rt
        table = list(params) + [None] * (drv.n_params - len(params))
        packed = drv.packed(rt, table, x.device)
        needs_grad = grad_on and any(p.requires_grad for p in params)
        if needs_grad and x.dtype == torch.uint8:
            raise NotImplementedError('uint8 frames are an encode-only input (use torch.no_grad(), or pass '
                                      'float32 frames in [0, 1] for training)')
        ws = drv.workspace(rt, x.shape[0], x.device, fresh=needs_grad)
        variational = drv.desc.n_heads == 2
        mu, logvar = drv.encode(x, table, packed, ws, variational)
        ctx.module, ctx.ws, ctx.packed, ctx.x, ctx.params = module, ws, packed, x, params
        ctx.variational = variational
        if variational:
            return mu, logvar
        return mu

    @staticmethod
    def backward(ctx, dmu, dlogvar=None):
        drv = ctx.module._driver
        params = ctx.params
        grads = [torch.zeros_like(p) if p.requires_grad else None for p in params]
        pad = [None] * (drv.n_params - len(params))
        dmu = None if dmu is None else dmu.contiguous()
        dlogvar = None if dlogvar is None else dlogvar.contiguous()
        drv.encode_bwd(ctx.x, dmu, dlogvar, list(params) + pad, ctx.packed, ctx.ws, grads + pad)
        return (None, None, None) + tuple(grads)


class DecodeFn(torch.autograd.Function):
    """Differentiable ConvAEDecoder.forward: (z, *decoder params) -> x_hat."""

    @staticmethod
    def forward(ctx, module, grad_on, z, *params):
        drv, rt = module._driver, module._rt
        head = [None] * (2 * drv.n_layers + 4)
        table = head + list(params)
        packed = drv.packed(rt, table, z.device)
        # a private workspace per graph that will be differentiated
        needs_grad = grad_on and (z.requires_grad or any(p.requires_grad for p in params))
        ws = drv.workspace(rt, z.shape[0], z.device, fresh=needs_grad)
        xhat = drv.decode(z, table, packed, ws, want_xhat=True)
        ctx.module, ctx.ws, ctx.packed, ctx.params, ctx.n = module, ws, packed, params, z.shape[0]
        ctx.z_needs_grad = z.requires_grad
        return xhat

    @staticmethod
    def backward(ctx, dxhat):
        drv = ctx.module._driver
        params = ctx.params
        head = [None] * (2 * drv.n_layers + 4)
        grads = [torch.zeros_like(p) if p.requires_grad else None for p in params]
        dz = drv.decode_bwd(ctx.n, dxhat.contiguous(), head + list(params), ctx.packed, ctx.ws,
                            head + grads, dxhat.device)
        return (None, None, dz if ctx.z_needs_grad else None) + tuple(grads)
