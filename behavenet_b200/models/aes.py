"""Convolutional autoencoder with the reference's object protocol, computed by sm_100a kernels.

Drop-in for ``behavenet.models.aes.{ConvAEEncoder, ConvAEDecoder, AE}`` (reference
``behavenet/models/aes.py:17-218, 221-488, 616-773``): same constructor (``hparams`` dict), same
``forward`` / ``loss`` / ``encoding`` / ``decoding`` signatures and tuple arities, same
``state_dict`` names and tensor layouts (checkpoints are interchangeable), same exceptions for
invalid hparams.  The ``nn.Conv2d`` / ``nn.ConvTranspose2d`` / ``nn.Linear`` submodules exist only
as parameter containers (so names, shapes and default initialisation are torch's own); their
``forward`` is never called -- all arithmetic runs in ``libbehavenet_b200.so`` and there is no
eager fallback: inputs on the CPU raise.
"""

import os

import numpy as np
import torch
from torch import nn

from behavenet_b200 import _lib, parallel
from behavenet_b200.models.base import BaseModule, BaseModel
from behavenet_b200.models._engine import CaeDriver, Runtime, EncodeFn, DecodeFn, output_padding

__all__ = ['ConvAEEncoder', 'ConvAEDecoder', 'LinearAEEncoder', 'LinearAEDecoder', 'AE', 'ConditionalAE', 'AEMSP', 'load_pretrained_ae']


_DP_TIMING = os.environ.get('BN_DP_TIMING', '0') == '1'
# BN_DP_TOP_BUCKET=1: encoder backward in two phases (bn_cae_encode_bwd_phase), the heads + top conv layer bucket
# (13 of the encoder's 17.5 MB) all-reduced under the layers below.  OFF by default -- measured on 8 x B200
# (profiles/r02_zz_dp_top_bucket_n8.txt): the collectives left exposed after the last backward kernel shrink from
# 0.12 to 0.09 ms, but the encoder's backward kernels slow down from 0.70 to 0.75 ms next to the second NCCL
# kernel (plus one more reduction launch): 2.556 ms per step against 2.503 ms without it.
_DP_TOP_BUCKET = os.environ.get('BN_DP_TOP_BUCKET', '0') == '1'
_DP_TIMES = []


def _dp_timing_report(ev):
    """BN_DP_TIMING=1 (diagnostic): device time of a data-parallel AE.loss call from the end of the forward pass --
    decoder backward | encoder backward | collectives still exposed after the last backward kernel; rank 0 prints the
    mean of every 10 calls."""
    _DP_TIMES.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
    if len(_DP_TIMES) % 10 == 0 and parallel.rank() == 0:
        import sys
        t = np.mean(np.array(_DP_TIMES[-10:]), axis=0)
        sys.stderr.write('[BN_DP_TIMING] decoder bwd %.3f ms, encoder bwd %.3f ms, exposed collectives %.3f ms\n' % tuple(t))


class ConvAEEncoder(BaseModule):
    """Convolutional encoder (reference aes.py:17-218)."""

    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self.encoder = None
        self.build_model()

    def __str__(self):
        s = 'Encoder architecture:\n'
        i = -1
        for i, module in enumerate(self.encoder):
            s += '    {:02d}: {}\n'.format(i, module)
        s += '    {:02d}: {}\n'.format(i + 1, self.FF)
        return s

    def build_model(self):
        """Parameter containers named like the reference's (aes.py:55-125)."""
        hp = self.hparams
        self._driver = CaeDriver(hp, 'encoder')     # raises NotImplementedError for unsupported variants
        self._rt = Runtime()
        self.encoder = nn.ModuleList()
        c_in = self._driver.img[0]          # frame channels (+ label images of a conditional encoder)
        for i, c_out in enumerate(hp['ae_encoding_n_channels']):
            x0, x1 = hp['ae_encoding_x_padding'][i]
            y0, y1 = hp['ae_encoding_y_padding'][i]
            if x0 == x1 and y0 == y1:
                padding = (y0, x0)
            else:
                self.encoder.add_module('zero_pad%i' % i, nn.ZeroPad2d((x0, x1, y0, y1)))
                padding = 0
            def conv(c_in=c_in, c_out=c_out, i=i, padding=padding):
                return nn.Conv2d(in_channels=c_in, out_channels=c_out, kernel_size=hp['ae_encoding_kernel_size'][i],
                                 stride=hp['ae_encoding_stride_size'][i], padding=padding)
            if hp.get('fit_sess_io_layers', False) and i == 0:
                # one input layer per session, chosen by ``dataset`` (aes.py:69-80)
                self.encoder.add_module('conv%i_sess_io_layers' % i, nn.ModuleList(
                    [conv() for _ in range(hp['n_datasets'])]))
            else:
                self.encoder.add_module('conv%i' % i, conv())
            self.encoder.add_module('relu%i' % i, nn.LeakyReLU(0.05))
            c_in = c_out
        last_conv_size = c_in * hp['ae_encoding_y_dim'][-1] * hp['ae_encoding_x_dim'][-1]
        self.FF = nn.Linear(last_conv_size, hp['n_ae_latents'])
        if hp.get('variational', False):
            self.logvar = nn.Linear(last_conv_size, hp['n_ae_latents'])

    def _conv_modules(self, dataset=None):
        """The conv layers of one pass: ``layer[dataset]`` for a per-session module list (aes.py:207)."""
        return [m[dataset] if isinstance(m, nn.ModuleList) else m for m in self.encoder
                if isinstance(m, (nn.Conv2d, nn.ModuleList))]

    def kernel_params(self, dataset=None):
        """Parameters in the order of the C parameter table (encoder side)."""
        ps = []
        for m in self._conv_modules(dataset):
            ps += [m.weight, m.bias]
        ps += [self.FF.weight, self.FF.bias]
        if self.hparams.get('variational', False):
            ps += [self.logvar.weight, self.logvar.bias]
        else:
            ps += [None, None]
        return ps

    def _heads(self, x, dataset=None):
        x = CaeDriver._check_input(x, 'encoder input', self._driver.img, allow_uint8=True)
        if x.requires_grad:
            raise NotImplementedError('gradients with respect to input frames are not computed')
        ps = self.kernel_params(dataset)
        if self.hparams.get('variational', False):
            return EncodeFn.apply(self, torch.is_grad_enabled(), x, *ps)
        return EncodeFn.apply(self, torch.is_grad_enabled(), x, *ps[:-2])

    def forward(self, x, dataset=None):
        """(z, pool_idx, output_size) -- or (mu, logvar, pool_idx, output_size) if variational
        (reference aes.py:181-218).  The pool lists are empty: there is no max-pooling path."""
        out = self._heads(x, dataset)
        if self.hparams.get('variational', False):
            return out[0], out[1], [], []
        return out, [], []


class ConvAEDecoder(BaseModule):
    """Convolutional decoder (reference aes.py:221-488)."""

    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self.decoder = None
        self.build_model()

    def __str__(self):
        s = 'Decoder architecture:\n'
        s += '    {:02d}: {}\n'.format(0, self.FF)
        for i, module in enumerate(self.decoder):
            s += '    {:02d}: {}\n'.format(i + 1, module)
        return s

    def build_model(self):
        hp = self.hparams
        if hp.get('ae_padding_type', 'same') not in ('same', 'valid'):
            raise ValueError('"%s" is not a valid padding type' % hp['ae_padding_type'])
        self._driver = CaeDriver(hp, 'decoder')
        self._rt = Runtime()
        c0, h0, w0 = hp['ae_decoding_starting_dim']
        self.FF = nn.Linear(hp['hidden_layer_size'], c0 * h0 * w0)
        self.decoder = nn.ModuleList()
        self.conv_t_pads = {}
        c_in = c0
        n = len(hp['ae_decoding_n_channels'])
        for i, c_out in enumerate(hp['ae_decoding_n_channels']):
            x0, x1 = hp['ae_decoding_x_padding'][i]
            y0, y1 = hp['ae_decoding_y_padding'][i]
            name = 'convtranspose%i' % i
            out_pad = 0
            if hp.get('ae_padding_type', 'same') == 'valid':
                padding, out_pad = (y0, x0), output_padding(hp, i)      # aes.py:382-405
                self.conv_t_pads[name] = None
            elif x0 == x1 and y0 == y1:
                padding = (y0, x0)
                self.conv_t_pads[name] = None
            else:
                padding = 0
                self.conv_t_pads[name] = [x0, x1, y0, y1]

            def convt(c_in=c_in, c_out=c_out, i=i, padding=padding, out_pad=out_pad):
                return nn.ConvTranspose2d(
                    in_channels=c_in, out_channels=c_out, kernel_size=(hp['ae_decoding_kernel_size'][i],) * 2,
                    stride=(hp['ae_decoding_stride_size'][i],) * 2, padding=padding, output_padding=out_pad)
            if hp.get('fit_sess_io_layers', False) and i == n - 1:
                # one output layer per session (aes.py:298-312)
                self.decoder.add_module(name + '_sess_io_layers', nn.ModuleList(
                    [convt() for _ in range(hp['n_datasets'])]))
                self.conv_t_pads[name + '_sess_io_layers'] = self.conv_t_pads[name]
            else:
                self.decoder.add_module(name, convt())
            if i == n - 1:
                self.decoder.add_module('sigmoid%i' % i, nn.Sigmoid())
            else:
                self.decoder.add_module('relu%i' % i, nn.LeakyReLU(0.05))
            c_in = c_out

    def _conv_modules(self, dataset=None):
        return [m[dataset] if isinstance(m, nn.ModuleList) else m for m in self.decoder
                if isinstance(m, (nn.ConvTranspose2d, nn.ModuleList))]

    def kernel_params(self, dataset=None):
        ps = [self.FF.weight, self.FF.bias]
        for m in self._conv_modules(dataset):
            ps += [m.weight, m.bias]
        return ps

    def forward(self, x, pool_idx=None, target_output_size=None, dataset=None):
        """x_hat of shape (n, C, H, W) (reference aes.py:432-488)."""
        x = CaeDriver._check_input(x, 'decoder input', (self._driver.L,))
        return DecodeFn.apply(self, torch.is_grad_enabled(), x, *self.kernel_params(dataset))


def _linear_input(x, what, n_features):
    """Frames flattened like the reference's ``x.view(x.size(0), -1)``."""
    if not x.is_cuda:
        raise RuntimeError('%s must live on a CUDA device (B200 kernels only; no CPU path)' % what)
    if x.dtype != torch.float32:
        raise TypeError('%s must be float32, got %s' % (what, x.dtype))
    x = x.contiguous().view(x.shape[0], -1)
    if x.shape[1] != n_features:
        raise ValueError('%s has %d values per frame, expected %d' % (what, x.shape[1], n_features))
    return x


class LinearAEEncoder(BaseModule):
    """Linear encoder (reference aes.py:491-547): ``z = x W^T + b`` on the flattened frame."""

    def __init__(self, n_latents, input_size):
        super().__init__()
        self.n_latents = n_latents
        self.input_size = input_size
        self.encoder = None
        self.decoder = None
        self.build_model()

    def __str__(self):
        return 'Encoder architecture:\n' + str('    {}\n'.format(self.encoder))

    def build_model(self):
        if not 1 <= self.n_latents <= 64:
            raise NotImplementedError('the linear-AE kernels hold 1..64 latents, got %d' % self.n_latents)
        self.encoder = nn.Linear(out_features=self.n_latents, in_features=int(np.prod(self.input_size)), bias=True)

    def forward(self, x, dataset=None):
        """(z, None, None) -- the Nones stand in for the conv encoder's pool lists (aes.py:530-547).  Inference
        only: training goes through ``AE.loss`` (one fused pass), there is no autograd bridge for this model."""
        P = int(np.prod(self.input_size))
        x = _linear_input(x, 'encoder input', P)
        z = torch.empty(x.shape[0], self.n_latents, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().bn_linae_forward(
            x.shape[0], P, self.n_latents, x.data_ptr(), self.encoder.weight.data_ptr(), self.encoder.bias.data_ptr(),
            None, z.data_ptr(), None, _lib.stream_ptr()), 'bn_linae_forward')
        return z, None, None


class LinearAEDecoder(BaseModule):
    """Linear decoder on the encoder's transposed weights plus its own bias (reference aes.py:550-613, the form
    ``AE.build_model`` constructs).  An independent weight matrix (``encoder=None``) has no kernel."""

    def __init__(self, n_latents, output_size, encoder=None):
        super().__init__()
        self.n_latents = n_latents
        self.output_size = output_size
        self.encoder = encoder          # registered as a submodule, like the reference (state_dict lists it twice)
        self.decoder = None
        self.build_model()

    def __str__(self):
        return 'Decoder architecture:\n    Encoder weights transposed (plus independent bias)\n'

    def build_model(self):
        if self.encoder is None:
            raise NotImplementedError('a linear decoder with its own weights has no B200 kernel; AE ties it to the encoder')
        self.bias = nn.Parameter(torch.zeros(int(np.prod(self.output_size))), requires_grad=True)

    def forward(self, x, dataset=None):
        """x_hat of shape (n, C, H, W) = z W + c (aes.py:588-613); inference only, see the encoder."""
        lin = self.encoder.encoder
        P = int(np.prod(self.output_size))
        z = _linear_input(x, 'decoder input', self.n_latents)
        n = z.shape[0]
        xhat = torch.empty((n,) + tuple(int(v) for v in self.output_size), dtype=torch.float32, device=z.device)
        _lib.check(_lib.lib().bn_linae_decode(
            n, P, self.n_latents, z.data_ptr(), lin.weight.data_ptr(), self.bias.data_ptr(), xhat.data_ptr(),
            _lib.stream_ptr()), 'bn_linae_decode')
        return xhat


class AE(BaseModel):
    """Convolutional (or linear) autoencoder (reference aes.py:616-773)."""

    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self.model_type = self.hparams['model_type']
        self.img_size = (
            self.hparams['n_input_channels'], self.hparams['y_pixels'], self.hparams['x_pixels'])
        self.encoding = None
        self.decoding = None
        # True: shard every batch over the ranks of behavenet_b200.parallel (one process per GPU)
        self.data_parallel = False
        self.build_model()

    def __str__(self):
        s = '\nAutoencoder architecture\n'
        s += '------------------------\n'
        s += self.encoding.__str__()
        s += self.decoding.__str__()
        return s + '\n'

    def build_model(self):
        self.hparams['hidden_layer_size'] = self.hparams['n_ae_latents']
        if self.model_type == 'conv':
            self.encoding = ConvAEEncoder(self.hparams)
            self.decoding = ConvAEDecoder(self.hparams)
        elif self.model_type == 'linear':
            if self.hparams.get('fit_sess_io_layers', False):
                raise NotImplementedError
            n_latents = self.hparams['n_ae_latents']
            self.encoding = LinearAEEncoder(n_latents, self.img_size)
            self.decoding = LinearAEDecoder(n_latents, self.img_size, self.encoding)
        else:
            raise ValueError('"%s" is an invalid model_type' % self.model_type)
        self._driver = getattr(self.encoding, '_driver', None)
        self._rt = Runtime()

    def forward(self, x, dataset=None, **kwargs):
        """(x_hat, z) (reference aes.py:695-720)."""
        if self.model_type == 'linear':
            z, _, _ = self.encoding(x)
            return self.decoding(z), z
        z, pool_idx, outsize = self.encoding(x, dataset=dataset)
        y = self.decoding(z, pool_idx, outsize, dataset=dataset)
        return y, z

    def _linear_loss(self, data, accumulate_grad, chunk_size):
        """AE.loss of the linear model (aes.py:722-773): three launches (encode, decode + masked squared error per
        reference chunk + its gradient, encoder weight gradient), gradients accumulated into ``.grad``."""
        enc, dec = self.encoding.encoder, self.decoding
        L, P = self.encoding.n_latents, int(np.prod(self.img_size))
        x = _linear_input(data['images'][0], "data['images'][0]", P)
        m = _linear_input(data['masks'][0].to(torch.float32), "data['masks'][0]", P) if 'masks' in data else None
        if 'shard' in data:
            beg, n_total = int(data['shard'][0]), int(data['shard'][1])
            xs, ms = x, m
        else:
            n_total = x.shape[0]
            beg, end = self._shard(n_total)
            xs, ms = x[beg:end], (None if m is None else m[beg:end])
        n = xs.shape[0]
        n_chunks = int(np.ceil(n_total / chunk_size))
        sse = torch.zeros(n_chunks, dtype=torch.float64, device=x.device)
        params = [enc.weight, enc.bias, dec.bias]
        if accumulate_grad:
            for p in params:
                if p.requires_grad and p.grad is None:
                    p.grad = torch.zeros_like(p)
        grads = [p.grad if (accumulate_grad and p.requires_grad) else None for p in params]
        if n > 0:
            lib = _lib.lib()
            ws = torch.empty(max(lib.bn_linae_workspace_bytes(n, P, L), 16), dtype=torch.uint8, device=x.device)
            _lib.check(lib.bn_linae_loss(
                n, P, L, xs.data_ptr(), _lib.ptr(ms), enc.weight.data_ptr(), enc.bias.data_ptr(), dec.bias.data_ptr(),
                int(chunk_size), beg, n_total, ws.data_ptr(), sse.data_ptr(), _lib.ptr(grads[0]), _lib.ptr(grads[1]),
                _lib.ptr(grads[2]), _lib.stream_ptr()), 'bn_linae_loss')
        if self.data_parallel and parallel.enabled():
            for g in grads:
                if g is not None:
                    parallel.all_reduce_sum(g)
            parallel.all_reduce_sum(sse)
        return {'loss': float(sse.sum().item()) / (float(P) * n_total)}

    # -- fused training step -----------------------------------------------------------------
    def invalidate_packed(self):
        """Force the next call to re-pack the GEMM-ordered weight copies (what an optimizer step does
        implicitly by bumping the parameter versions)."""
        self._rt.packed_key = None

    def _kernel_params(self, dataset=None):
        """The C parameter table of one pass; ``dataset`` picks the session's input / output layers when the
        model has them (``fit_sess_io_layers``)."""
        return self.encoding.kernel_params(dataset) + self.decoding.kernel_params(dataset)

    def _extra_trainable(self):
        """Trainable parameters that are not in the C parameter table (PS-VAE label head)."""
        return []

    def _grad_table(self, params):
        """.grad tensors to accumulate into (autograd semantics: created as zeros if absent).
        Freshly created gradients are views of one flat buffer so that a data-parallel run needs
        a single all-reduce."""
        everything = [p for p in params if p is not None] + self._extra_trainable()
        missing = [p for p in everything if p.requires_grad and p.grad is None]
        if missing:
            total = sum(p.numel() for p in missing)
            flat = torch.zeros(total, dtype=torch.float32, device=missing[0].device)
            o = 0
            for p in missing:
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
            self._rt.bufs['flat_grad'] = flat
            self._rt.bufs['flat_grad_ptrs'] = [p.grad.data_ptr() for p in missing]
            # the decoder's parameters form one contiguous bucket of the flat buffer (they are final after
            # bn_cae_decode_bwd, before the encoder's backward pass starts)
            dec = {id(p) for p in self.decoding.parameters()}
            o, lo, hi, cnt = 0, None, None, 0
            for p in missing:
                if id(p) in dec:
                    lo = o if lo is None else lo
                    hi = o + p.numel()
                    cnt += p.numel()
                o += p.numel()
            self._rt.bufs['flat_grad_dec'] = (lo, hi) if lo is not None and hi - lo == cnt else None
            # the encoder's heads and top conv layer form the contiguous bucket right below the decoder's: they
            # are final after phase 1 of bn_cae_encode_bwd_phase
            n_enc = 2 * self._driver.n_layers
            top = {id(p) for p in params[n_enc - 2:n_enc + 4] if p is not None}
            o, lo, hi, cnt = 0, None, None, 0
            for p in missing:
                if id(p) in top:
                    lo = o if lo is None else lo
                    hi = o + p.numel()
                    cnt += p.numel()
                o += p.numel()
            self._rt.bufs['flat_grad_top'] = (lo, hi) if lo is not None and hi - lo == cnt else None
        return [None if (p is None or not p.requires_grad) else p.grad for p in params]

    def _shard(self, n):
        """Contiguous frame range of this rank (data-parallel) -> (begin, end)."""
        if not (self.data_parallel and parallel.enabled()):
            return 0, n
        return parallel.shard_range(n)

    def _flat_is_live(self, params):
        flat = self._rt.bufs.get('flat_grad')
        ptrs = self._rt.bufs.get('flat_grad_ptrs')
        everything = [p for p in params if p is not None] + self._extra_trainable()
        live = [p.grad.data_ptr() for p in everything if p.requires_grad]
        return flat is not None and ptrs == live, everything

    def _allreduce_begin(self, params, which='dec'):
        """Start the all-reduce of the decoder-side gradient bucket as soon as the decoder's backward pass
        is enqueued: the collective runs on the communication stream underneath the encoder's backward
        kernels ("decoder-side buckets overlap with encoder backward").  Returns a
        handle for ``_allreduce`` or None when the gradients are not views of the flat buffer."""
        ok, _ = self._flat_is_live(params)
        span = self._rt.bufs.get('flat_grad_' + which)
        if not ok or span is None or not parallel.overlap_enabled():
            return None
        flat = self._rt.bufs['flat_grad']
        return (parallel.all_reduce_sum_async(flat[span[0]:span[1]]), span)

    def _allreduce(self, params, extra, pending=None):
        """All-reduce(SUM) of the flat gradient -- minus the buckets already in flight (``pending``: one
        ``_allreduce_begin`` handle or a list of them) -- and of the loss partial sums."""
        if pending is not None and not isinstance(pending, list):
            pending = [pending]
        pending = [p for p in (pending or []) if p is not None]
        ok, everything = self._flat_is_live(params)
        if ok:
            flat = self._rt.bufs['flat_grad']
            o = 0
            for lo, hi in sorted(span for _, span in pending) + [(flat.numel(), flat.numel())]:
                if lo > o:
                    parallel.all_reduce_sum(flat[o:lo])
                o = max(o, hi)
        else:
            if pending:
                for work, _ in pending:
                    work.wait()
                raise RuntimeError('gradient buffers changed between the bucketed all-reduce and its completion')
            for p in everything:
                if p.requires_grad:
                    parallel.all_reduce_sum(p.grad)
        if extra is not None:
            parallel.all_reduce_sum(extra)
        for work, _ in pending:
            work.wait()

    def loss(self, data, dataset=0, accumulate_grad=True, chunk_size=200):
        """MSE loss (+ gradients) with the reference's chunk semantics (aes.py:722-773): the batch
        is treated as ceil(n / chunk_size) chunks, each contributing the gradient of its own mean
        squared error; the returned value is the frame-weighted mean of the chunk losses.

        One fused pass: encoder kernels -> decoder kernels with the squared error and its
        gradient computed in the last layer's epilogue -> backward kernels accumulating into
        ``.grad``.  A single device->host read of the per-chunk sums ends the call.
        """
        if self.model_type == 'linear':
            return self._linear_loss(data, accumulate_grad, chunk_size)
        x = data['images'][0]
        m = data['masks'][0] if 'masks' in data else None
        drv, rt = self._driver, self._rt
        x = drv._check_input(x, "data['images'][0]", drv.img)
        if m is not None:
            m = drv._check_input(m.to(torch.float32), "data['masks'][0]", drv.img)
        if 'shard' in data:
            # data-parallel callers that stage only their own frames: data['shard'] = (first frame,
            # frames in the whole batch); chunk membership still follows the whole batch
            beg, n_total = int(data['shard'][0]), int(data['shard'][1])
            end = beg + x.shape[0]
            local = True
        else:
            n_total = x.shape[0]
            beg, end = self._shard(n_total)
            local = False
        n_chunks = int(np.ceil(n_total / chunk_size))
        n = end - beg
        params = self._kernel_params(dataset)
        device = x.device
        sse = torch.zeros(n_chunks, dtype=torch.float64, device=device)
        pending = None
        if n > 0:
            xs = x if local else x[beg:end]
            ms = None if m is None else (m if local else m[beg:end])
            packed = drv.packed(rt, params, device)
            ws = drv.workspace(rt, n, device)
            z, _ = drv.encode(xs, params, packed, ws, False)
            numel = float(np.prod(drv.img))
            drv.decode(z, params, packed, ws, want_xhat=False, target=xs, mask=ms,
                       chunk_size=chunk_size, frame_offset=beg, n_total=n_total,
                       grad_coef=2.0 / numel, sse=sse)
        dp = self.data_parallel and parallel.enabled()
        # the per-chunk loss sums are final after the forward pass: their (tiny, latency-bound) all-reduce is started
        # here and completes under the backward pass instead of after it
        sse_work = parallel.all_reduce_sum_async(sse) if (dp and parallel.overlap_enabled()) else None
        timing = dp and _DP_TIMING
        if timing:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        if n > 0 and accumulate_grad:
            grads = self._grad_table(params)
            dz = drv.decode_bwd(n, None, params, packed, ws, grads, device)
            if dp:
                pending = [self._allreduce_begin(params)]
            if timing:
                ev[1].record()
            if dp and _DP_TOP_BUCKET and pending[0] is not None and self._rt.bufs.get('flat_grad_top') is not None:
                # heads + top conv layer first; their bucket is all-reduced under the layers below
                drv.encode_bwd(xs, dz, None, params, packed, ws, grads, phase=1)
                pending.append(self._allreduce_begin(params, 'top'))
                drv.encode_bwd(xs, dz, None, params, packed, ws, grads, phase=2)
            else:
                drv.encode_bwd(xs, dz, None, params, packed, ws, grads)
        elif accumulate_grad:
            self._grad_table(params)
            if dp:      # a rank without frames issues the same collectives
                pending = [self._allreduce_begin(params)]
                if _DP_TOP_BUCKET and pending[0] is not None and self._rt.bufs.get('flat_grad_top') is not None:
                    pending.append(self._allreduce_begin(params, 'top'))
        if timing:
            if not (n > 0 and accumulate_grad):
                ev[1].record()
            ev[2].record()
        if dp:
            if accumulate_grad:
                self._allreduce(params, None if sse_work is not None else sse, pending)
            elif sse_work is None:
                parallel.all_reduce_sum(sse)
            if sse_work is not None:
                sse_work.wait()
        if timing:
            ev[3].record()
            ev[3].synchronize()
            _dp_timing_report(ev)
        numel = float(np.prod(drv.img))
        loss_val = float(sse.sum().item()) / (numel * n_total)
        return {'loss': loss_val}


from behavenet_b200.fitting.losses import mse as _masked_mse      # noqa: E402  (losses.mse, losses.py:36-59)


class _AutogradChunkedAE(AE):
    """AE variants whose loss has terms outside the fused encode->decode->loss pass: the conv stacks run in
    the same kernels through ``EncodeFn`` / ``DecodeFn`` and the few small terms in between are ordinary
    torch ops on (n, latents) tensors, chunk by chunk like the reference."""

    def invalidate_packed(self):
        self.encoding._rt.packed_key = None
        self.decoding._rt.packed_key = None

    def _chunk_loop(self, data, chunk_size, accumulate_grad, chunk_loss, keys):
        if (self.data_parallel and parallel.enabled()) or 'shard' in data:
            raise NotImplementedError('%s has no data-parallel path (no frame sharding, no gradient all-reduce); '
                                      'run it with one process per model' % type(self).__name__)
        x = data['images'][0]
        n = x.shape[0]
        vals = {k: 0.0 for k in keys}
        for b in range(0, n, chunk_size):
            e = min(b + chunk_size, n)
            terms = chunk_loss(b, e)
            if accumulate_grad:
                terms[keys[0]].backward()
            for k in keys:
                vals[k] += terms[k].item() * (e - b)
        return {k: v / n for k, v in vals.items()}


class ConditionalAE(_AutogradChunkedAE):
    """Conditional autoencoder (reference aes.py:776-903): labels are concatenated to the latents in front
    of the decoder's FF layer; with ``conditional_encoder`` their one-hot images are extra input channels."""

    def __init__(self, hparams):
        if hparams['model_type'] == 'linear':
            raise NotImplementedError
        super().__init__(hparams)

    def build_model(self):
        self.hparams['hidden_layer_size'] = self.hparams['n_ae_latents'] + self.hparams['n_labels']
        self.encoding = ConvAEEncoder(self.hparams)
        self.decoding = ConvAEDecoder(self.hparams)
        self._driver = self.encoding._driver
        self._rt = Runtime()

    def forward(self, x, dataset=None, labels=None, labels_2d=None, **kwargs):
        """(x_hat, z) (aes.py:810-836)."""
        if self.hparams['conditional_encoder']:
            x = torch.cat((x, labels_2d), dim=1)
        z, pool_idx, outsize = self.encoding(x, dataset=dataset)
        y = self.decoding(torch.cat((z, labels), dim=1), pool_idx, outsize, dataset=dataset)
        return y, z

    def loss(self, data, dataset=0, accumulate_grad=True, chunk_size=200):
        """{'loss'}: masked MSE per chunk (aes.py:838-903)."""
        x, y = data['images'][0], data['labels'][0]
        m = data['masks'][0] if 'masks' in data else None
        y2d = data['labels_sc'][0] if self.hparams['conditional_encoder'] else None

        def chunk_loss(b, e):
            x_hat, _ = self.forward(x[b:e], labels=y[b:e], labels_2d=None if y2d is None else y2d[b:e],
                                    dataset=dataset)
            return {'loss': _masked_mse(x[b:e], x_hat, None if m is None else m[b:e])}
        return self._chunk_loop(data, chunk_size, accumulate_grad, chunk_loss, ['loss'])


class AEMSP(_AutogradChunkedAE):
    """Autoencoder with matrix subspace projection (reference aes.py:906-1217): a bias-free linear map of
    the latents predicts the labels; ``U`` completes it to an orthogonal basis when the model is saved."""

    def __init__(self, hparams):
        if hparams['model_type'] == 'linear':
            raise NotImplementedError
        if hparams['n_ae_latents'] < hparams['n_labels']:
            raise ValueError('AEMSP model must contain at least as many latents as labels')
        self.n_latents = hparams['n_ae_latents']
        self.n_labels = hparams['n_labels']
        super().__init__(hparams)

    def build_model(self):
        self.hparams['hidden_layer_size'] = self.hparams['n_ae_latents']
        self.encoding = ConvAEEncoder(self.hparams)
        self.decoding = ConvAEDecoder(self.hparams)
        self._driver = self.encoding._driver
        self._rt = Runtime()
        self.projection = nn.Linear(self.n_latents, self.n_labels, bias=False)
        with torch.no_grad():
            self.U = nn.Linear(self.n_latents, self.n_latents, bias=False)

    def forward(self, x, dataset=None, **kwargs):
        """(x_hat, z, y) (aes.py:973-995)."""
        z, pool_idx, outsize = self.encoding(x, dataset=dataset)
        y = self.projection(z)
        x_hat = self.decoding(z, pool_idx, outsize, dataset=dataset)
        return x_hat, z, y

    def loss(self, data, dataset=0, accumulate_grad=True, chunk_size=200):
        """{'loss', 'loss_mse', 'loss_msp', 'labels_r2'} (aes.py:997-1077)."""
        from ..fitting.losses import r2_variance_weighted
        x, y = data['images'][0], data['labels'][0]
        m = data['masks'][0] if 'masks' in data else None
        y_hats = []

        def chunk_loss(b, e):
            x_hat, z, y_hat = self.forward(x[b:e], dataset=dataset)
            l_mse = _masked_mse(x[b:e], x_hat, None if m is None else m[b:e])
            l_msp = _masked_mse(y[b:e], y_hat) + _masked_mse(z, torch.matmul(y_hat, self.projection.weight))
            y_hats.append(y_hat.detach())
            return {'loss': l_mse + self.hparams['msp.alpha'] * l_msp, 'loss_mse': l_mse, 'loss_msp': l_msp}
        out = self._chunk_loop(data, chunk_size, accumulate_grad, chunk_loss, ['loss', 'loss_mse', 'loss_msp'])
        out['labels_r2'] = r2_variance_weighted(y.detach().cpu().numpy(), torch.cat(y_hats, 0).cpu().numpy())
        return out

    def save(self, filepath):
        self.create_orthogonal_matrix()
        super().save(filepath)

    def create_orthogonal_matrix(self):
        """U = [projection; null-space basis of the projection] (aes.py:1083-1097)."""
        from scipy.linalg import null_space
        M = self.projection.weight.data.detach().cpu().numpy()
        U = np.concatenate([M, null_space(M).T], axis=0)
        dev = self.projection.weight.device
        with torch.no_grad():
            self.U.weight = nn.Parameter(torch.from_numpy(U).float().to(dev), requires_grad=False)

    def get_transformed_latents(self, inputs, dataset=None, as_numpy=True):
        """Latents in the [labels | rest] basis from frames (4-D input) or latents (2-D) (aes.py:1099-1138)."""
        if not isinstance(inputs, torch.Tensor):
            inputs = torch.Tensor(inputs)
        inputs = inputs.to(self.U.weight.device)
        z = inputs if inputs.dim() == 2 else self.encoding(inputs, dataset=dataset)[0]
        out = self.U(z)
        return out.cpu().detach().numpy() if as_numpy else out

    def get_inverse_transformed_latents(self, latents, as_numpy=True):
        """Back to the encoder's basis (aes.py:1140-1163)."""
        if not isinstance(latents, torch.Tensor):
            latents = torch.Tensor(latents)
        out = torch.matmul(latents.to(self.U.weight.device), self.U.weight)
        return out.cpu().detach().numpy() if as_numpy else out

    def sample(self, x=None, dataset=None, latents=None, labels=None, labels_2d=None):
        """Decode user-chosen labels and/or transformed latents, the rest taken from ``x`` (aes.py:1165-1217)."""
        if latents is None or labels is None:
            tr = self.get_transformed_latents(x, dataset)
        else:
            tr = np.full((latents.shape[0], self.n_latents), np.nan)
        if labels is not None:
            tr[:, :self.n_labels] = labels
        if latents is not None:
            tr[:, self.n_labels:] = latents
        z = self.get_inverse_transformed_latents(torch.from_numpy(np.asarray(tr)).float(), as_numpy=False)
        return self.decoding(z.contiguous(), None, None, dataset=dataset)


def load_pretrained_ae(model, hparams):
    """Load pretrained weights into ``model`` (reference aes.py:1220-1274): FF layers are dropped
    when the latent count differs, everything else must match."""
    import os
    if hparams.get('pretrained_weights_path') is None or hparams['pretrained_weights_path'] == '':
        return model
    path = hparams['pretrained_weights_path']
    if not os.path.exists(path):
        raise FileNotFoundError('pretrained weights %s do not exist' % path)
    loaded = torch.load(path, map_location=lambda storage, loc: storage)
    own = model.state_dict()
    if loaded['encoding.FF.weight'].shape == own['encoding.FF.weight'].shape:
        model.load_state_dict(loaded, strict=False)
    else:
        print('warning: number of latents differ; pretrained FF layers are skipped')
        for k in ['encoding.FF.weight', 'encoding.FF.bias', 'decoding.FF.weight', 'decoding.FF.bias']:
            loaded.pop(k, None)
        model.load_state_dict(loaded, strict=False)
    return model
