"""CPU oracle for hot path 1 (convolutional autoencoder / PS-VAE).  TEST INFRASTRUCTURE ONLY.

This is a plain-PyTorch fp32 (or fp64) restatement of what the reference computes for the CAE
path, written functionally over a reference-named ``state_dict``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
it; the product (``behavenet_b200``) never does.

Pinning: the restatement is checked against the UNMODIFIED reference classes
(``behavenet.models.AE`` / ``PSVAE`` imported from /root/reference) by
``oracle/gen_golden.py`` -- which also writes the committed fixtures under ``tests/golden/`` --
and against the closed forms of the reference's own ``tests/test_fitting/test_losses.py``
(re-stated in ``tests/test_oracle_cae.py``).

Reference lines followed (all under /root/reference/behavenet):
  encoder forward   models/aes.py:181-218   (ZeroPad2d -> Conv2d -> LeakyReLU(0.05), flatten C,H,W)
  decoder forward   models/aes.py:432-488   (FF -> view(C0,H0,W0) -> ConvT -> crop -> LeakyReLU .. Sigmoid)
  AE.loss           models/aes.py:722-773   (<=200-frame chunks, per-chunk mean loss + backward)
  PS encoder        models/vaes.py:1319-1363
  PSVAE.forward     models/vaes.py:571-601 ; reparameterize 17-35 (std = exp(logvar))
  PSVAE.loss        models/vaes.py:603-729
  losses            fitting/losses.py:36-59 (mse) 62-96 (gaussian_ll) 130-147 (kl) 284-372 (decomposed kl)
"""

import math

import numpy as np
import torch
import torch.nn.functional as F

LN2PI = math.log(2 * math.pi)
LEAK = 0.05


# ------------------------------------------------------------------------------------------------
# losses (fitting/losses.py)
# ------------------------------------------------------------------------------------------------

def mse(y_pred, y_true, masks=None):
    """losses.py:36-59 -- mean over ALL elements of (diff^2 * mask)."""
    d = (y_pred - y_true) ** 2
    if masks is not None:
        d = d * masks
    return d.mean()


def gaussian_ll(y_pred, y_mean, masks=None, std=1):
    """losses.py:62-96 -- sum over dims, mean over batch, fixed std."""
    n_dims = int(np.prod(y_pred.shape[1:]))
    d = (y_pred - y_mean) ** 2
    if masks is not None:
        d = d * masks
    per_frame = d.reshape(d.shape[0], -1).sum(1)
    ll = -(0.5 * LN2PI + 0.5 * math.log(std ** 2)) * n_dims - (0.5 / std ** 2) * per_frame
    return ll.mean()


def gaussian_ll_to_mse(ll, n_dims, gaussian_std=1, mse_std=1):
    """losses.py:99-127."""
    v = float(ll) + (0.5 * LN2PI + 0.5 * math.log(gaussian_std ** 2)) * n_dims
    v *= -(gaussian_std ** 2) / 0.5
    v /= n_dims
    return v / (mse_std ** 2)


def kl_div_to_std_normal(mu, logvar):
    """losses.py:130-147."""
    return (0.5 * (logvar.exp() - logvar + mu ** 2 - 1).sum(1)).mean()


def decomposed_kl(z, mu, logvar):
    """losses.py:284-351 -- minibatch estimators (index-code MI, total correlation, dim-wise KL)."""
    # [j, i, l]: log q(z_j,l | x_i)
    lq = -0.5 * (torch.exp(-logvar)[None] * (z[:, None] - mu[None]) ** 2 + logvar[None] + LN2PI)
    joint = lq.sum(2)                                   # (j, i)
    log_qz = torch.logsumexp(joint, dim=1)              # log sum_i prod_l
    log_qz_diag = torch.diagonal(joint)                 # log q(z_j | x_j)
    log_qz_prod = torch.logsumexp(lq, dim=1).sum(1)     # sum_l log sum_i
    log_pz = (-0.5 * (z ** 2 + LN2PI)).sum(1)
    return ((log_qz_diag - log_qz).mean(), (log_qz - log_qz_prod).mean(),
            (log_qz_prod - log_pz).mean())


# ------------------------------------------------------------------------------------------------
# encoder / decoder (models/aes.py)
# ------------------------------------------------------------------------------------------------

def n_conv_layers(hparams):
    return len(hparams['ae_encoding_n_channels'])


# ---- optional emulation of the TF32 tensor-core arithmetic of the benchmarked path ----------------
# The product's default compute mode multiplies TF32 operands with fp32 accumulation in the fat conv
# layers: weights are rounded to TF32 once (cvt.rna: nearest, ties away from zero) when they are packed
# into GEMM order; activations / upstream gradients are consumed as stored, the tensor core reading
# only the top 19 bits (truncation).  Under ``tf32_emulation()`` the conv / transposed-conv calls below
# reproduce exactly that operand rounding in the forward, backward-data and weight-gradient products
# (everything else -- biases, activations, FF layers, accumulation -- stays fp32/fp64), so the
# tensor-core path can be pinned to round-off tolerances instead of to the 1e-2-level distance between
# TF32 and fp32 arithmetic.  The first encoder layer's weight gradient and the last decoder layer's
# forward + weight gradient run on fp32 CUDA-core kernels in the product and stay exact here
# (behavenet_b200/csrc/cae_thin.cu).
_EMU = {'on': False, 'act': 'trunc'}


class tf32_emulation:
    def __init__(self, act='trunc'):
        self.act = act

    def __enter__(self):
        self.prev = dict(_EMU)
        _EMU.update(on=True, act=self.act)

    def __exit__(self, *exc):
        _EMU.update(self.prev)


def tf32_rna(t):
    """cvt.rna.tf32.f32: round to 10 mantissa bits, nearest, ties away from zero."""
    f = t.detach().to(torch.float32).contiguous()
    b = f.view(torch.int32)
    b = (b + 0x1000) & ~0x1FFF
    return b.view(torch.float32).to(t.dtype)


def tf32_trunc(t):
    """What the tensor core reads of an fp32 operand: the low 13 mantissa bits are ignored."""
    f = t.detach().to(torch.float32).contiguous()
    return (f.view(torch.int32) & ~0x1FFF).view(torch.float32).to(t.dtype)


def _q_act(t):
    return tf32_trunc(t) if _EMU['act'] == 'trunc' else tf32_rna(t)


class _EmuConv(torch.autograd.Function):
    """conv2d / conv_transpose2d (stride s, no padding) with TF32 operand rounding per product."""

    @staticmethod
    def forward(ctx, x, w, b, s, transposed, q_fwd, q_dgrad, q_wgrad):
        op = F.conv_transpose2d if transposed else F.conv2d
        ctx.save_for_backward(x, w)
        ctx.cfg = (s, transposed, q_dgrad, q_wgrad)
        xin, win = (_q_act(x), tf32_rna(w)) if q_fwd else (x.detach(), w.detach())
        return op(xin, win, b.detach(), stride=s)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        s, transposed, q_dgrad, q_wgrad = ctx.cfg
        op = F.conv_transpose2d if transposed else F.conv2d
        dx = dw = None
        with torch.enable_grad():
            if ctx.needs_input_grad[0]:
                xin = x.detach().requires_grad_(True)
                y = op(xin, tf32_rna(w) if q_dgrad else w.detach(), None, stride=s)
                dx, = torch.autograd.grad(y, xin, _q_act(dy) if q_dgrad else dy)
            if ctx.needs_input_grad[1]:
                win = w.detach().requires_grad_(True)
                y = op(_q_act(x) if q_wgrad else x.detach(), win, None, stride=s)
                dw, = torch.autograd.grad(y, win, _q_act(dy) if q_wgrad else dy)
        db = dy.sum((0, 2, 3)) if ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None, None, None, None


def _conv(x, w, b, s, layer, n_layers, transposed):
    if not _EMU['on']:
        return (F.conv_transpose2d if transposed else F.conv2d)(x, w, b, stride=s)
    if transposed:      # last decoder layer: fp32 forward and weight gradient, TF32 backward-data
        last = layer == n_layers - 1
        return _EmuConv.apply(x, w, b, s, True, not last, True, not last)
    return _EmuConv.apply(x, w, b, s, False, True, True, layer > 0)


def _io_name(hparams, stem, i, n, dataset, first):
    """Parameter-name stem of conv layer i: the first encoder layer / last decoder layer are per-session
    module lists when ``fit_sess_io_layers`` is set (aes.py:69-80, 298-312) -- not the decoder's when it
    ends in a feed-forward layer."""
    io = hparams.get('fit_sess_io_layers', False) and (i == 0 if first else (
        i == n - 1 and not hparams.get('ae_decoding_last_FF_layer', False)))
    return '%s%i_sess_io_layers.%i' % (stem, i, dataset) if io else '%s%i' % (stem, i)


def encoder_features(sd, hparams, x, prefix='encoding.', dataset=None):
    """Conv stack of ConvAEEncoder.forward (aes.py:181-214); returns flattened (N, C*H*W)."""
    n = n_conv_layers(hparams)
    for i in range(n):
        x0, x1 = hparams['ae_encoding_x_padding'][i]
        y0, y1 = hparams['ae_encoding_y_padding'][i]
        s = hparams['ae_encoding_stride_size'][i]
        # symmetric: Conv2d(padding=(y0,x0)); asymmetric: ZeroPad2d((x0,x1,y0,y1)) + padding 0
        # (aes.py:141-155) -- numerically the same explicit zero pad
        x = F.pad(x, (x0, x1, y0, y1))
        name = prefix + 'encoder.' + _io_name(hparams, 'conv', i, n, dataset, True)
        x = _conv(x, sd[name + '.weight'], sd[name + '.bias'], s, i, n, False)
        x = F.leaky_relu(x, LEAK)
    return x.reshape(x.shape[0], -1)


def encode(sd, hparams, x, prefix='encoding.', dataset=None):
    """ConvAEEncoder.forward -> z   (or (mu, logvar) if hparams['variational'])."""
    h = encoder_features(sd, hparams, x, prefix, dataset)
    z = F.linear(h, sd[prefix + 'FF.weight'], sd[prefix + 'FF.bias'])
    if hparams.get('variational', False):
        return z, F.linear(h, sd[prefix + 'logvar.weight'], sd[prefix + 'logvar.bias'])
    return z


def decode(sd, hparams, z, prefix='decoding.', dataset=None):
    """ConvAEDecoder.forward (aes.py:432-488)."""
    c0, h0, w0 = hparams['ae_decoding_starting_dim']
    x = F.linear(z, sd[prefix + 'FF.weight'], sd[prefix + 'FF.bias']).view(-1, c0, h0, w0)
    n = len(hparams['ae_decoding_n_channels'])
    valid = hparams.get('ae_padding_type', 'same') == 'valid'
    for i in range(n):
        x0, x1 = hparams['ae_decoding_x_padding'][i]
        y0, y1 = hparams['ae_decoding_y_padding'][i]
        s = hparams['ae_decoding_stride_size'][i]
        name = prefix + 'decoder.' + _io_name(hparams, 'convtranspose', i, n, dataset, False)
        if valid:
            # 'valid' (aes.py:382-405): ConvTranspose2d(padding=(y0, x0), output_padding=target - full size);
            # the rows / columns that output_padding appends receive no input, only the bias
            x = _conv(x, sd[name + '.weight'], torch.zeros_like(sd[name + '.bias']), s, i, n, True)
            x = x[:, :, y0:x.shape[2] - y0, x0:x.shape[3] - x0]
            opy = hparams['ae_decoding_y_dim'][i] - x.shape[2]
            opx = hparams['ae_decoding_x_dim'][i] - x.shape[3]
            x = F.pad(x, (0, opx, 0, opy)) + sd[name + '.bias'].view(1, -1, 1, 1)
        else:
            # symmetric pads go in as ConvTranspose2d(padding=...), asymmetric as a crop afterwards
            # (aes.py:404-418, 467-470): both are "full transposed conv, then crop"
            x = _conv(x, sd[name + '.weight'], sd[name + '.bias'], s, i, n, True)
            x = x[:, :, y0:x.shape[2] - y1, x0:x.shape[3] - x1]
        x = torch.sigmoid(x) if i == n - 1 else F.leaky_relu(x, LEAK)
    return x


def ae_forward(sd, hparams, x, dataset=None):
    """AE.forward (aes.py:695-720) -> (x_hat, z)."""
    z = encode(sd, hparams, x, dataset=dataset)
    return decode(sd, hparams, z, dataset=dataset), z


# ---- linear autoencoder (models/aes.py:491-613; AE.build_model ties the decoder to the encoder, aes.py:684-687) ----

def make_linear_hparams(n_input_channels, y_pixels, x_pixels, n_ae_latents):
    """The keys AE reads for ``model_type='linear'`` (aes.py:660-687)."""
    return {'model_class': 'ae', 'model_type': 'linear', 'n_input_channels': n_input_channels, 'y_pixels': y_pixels,
            'x_pixels': x_pixels, 'n_ae_latents': n_ae_latents, 'fit_sess_io_layers': False}


def init_linear_state_dict(hparams, seed=0, dtype=torch.float32):
    """Random parameters under the reference's names.  The decoder registers the encoder module as a submodule
    (aes.py:573), so its tensors appear a second time as ``decoding.encoder.encoder.*``."""
    g = torch.Generator().manual_seed(seed)
    P = hparams['n_input_channels'] * hparams['y_pixels'] * hparams['x_pixels']
    L = hparams['n_ae_latents']
    bound = 1 / math.sqrt(P)
    sd = {'encoding.encoder.weight': ((torch.rand(L, P, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype),
          'encoding.encoder.bias': ((torch.rand(L, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype),
          'decoding.bias': (torch.rand(P, generator=g, dtype=torch.float64) * 0.5 + 0.25).to(dtype)}
    sd['decoding.encoder.encoder.weight'] = sd['encoding.encoder.weight']
    sd['decoding.encoder.encoder.bias'] = sd['encoding.encoder.bias']
    return sd


def linear_ae_forward(sd, hparams, x):
    """AE.forward, linear branch (aes.py:714-716) -> (x_hat, z)."""
    W, b, c = sd['encoding.encoder.weight'], sd['encoding.encoder.bias'], sd['decoding.bias']
    z = F.linear(x.reshape(x.shape[0], -1), W, b)
    return (F.linear(z, W.t()) + c).view(x.shape), z


def linear_ae_loss(sd, hparams, x, masks=None, chunk_size=200, want_grads=True):
    """AE.loss (aes.py:722-773) of the linear model; gradients under the three distinct parameter names."""
    names = ['encoding.encoder.weight', 'encoding.encoder.bias', 'decoding.bias']
    params = {k: sd[k].detach().clone().requires_grad_(want_grads) for k in names}
    total = 0.0
    for b, e in _chunks(x.shape[0], chunk_size):
        x_hat, _ = linear_ae_forward(params, hparams, x[b:e])
        loss = mse(x[b:e], x_hat, None if masks is None else masks[b:e])
        if want_grads:
            loss.backward()
        total += loss.item() * (e - b)
    grads = {k: v.grad for k, v in params.items()} if want_grads else {}
    return {'loss': total / x.shape[0]}, grads


def _chunks(n, chunk_size):
    return [(b, min(b + chunk_size, n)) for b in range(0, n, chunk_size)]


def ae_loss(sd, hparams, x, masks=None, chunk_size=200, want_grads=True, dataset=None):
    """AE.loss (aes.py:722-773): returns ({'loss': float}, {name: grad}) with the reference's
    chunk semantics (each chunk back-propagates its own mean)."""
    params = {k: v.detach().clone().requires_grad_(want_grads) for k, v in sd.items()}
    total = 0.0
    for b, e in _chunks(x.shape[0], chunk_size):
        x_in = x[b:e]
        x_hat, _ = ae_forward(params, hparams, x_in, dataset)
        loss = mse(x_in, x_hat, None if masks is None else masks[b:e])
        if want_grads:
            loss.backward()
        total += loss.item() * (e - b)
    grads = {k: v.grad for k, v in params.items() if v.grad is not None} if want_grads else {}
    return {'loss': total / x.shape[0]}, grads


# ------------------------------------------------------------------------------------------------
# Conditional AE / AE with matrix subspace projection (models/aes.py:776-1217)
# ------------------------------------------------------------------------------------------------

def cond_ae_forward(sd, hparams, x, labels, labels_2d=None):
    """ConditionalAE.forward (aes.py:810-836): labels join the latents before the decoder's FF layer; with
    ``conditional_encoder`` their one-hot images join the frames' channels."""
    if hparams.get('conditional_encoder', False):
        x = torch.cat((x, labels_2d), dim=1)
    z = encode(sd, hparams, x)
    return decode(sd, hparams, torch.cat((z, labels), dim=1)), z


def cond_ae_loss(sd, hparams, x, labels, labels_2d=None, masks=None, chunk_size=200, want_grads=True):
    """ConditionalAE.loss (aes.py:838-903)."""
    params = {k: v.detach().clone().requires_grad_(want_grads) for k, v in sd.items()}
    total = 0.0
    for b, e in _chunks(x.shape[0], chunk_size):
        x_in = x[b:e]
        x_hat, _ = cond_ae_forward(params, hparams, x_in, labels[b:e],
                                   None if labels_2d is None else labels_2d[b:e])
        loss = mse(x_in, x_hat, None if masks is None else masks[b:e])
        if want_grads:
            loss.backward()
        total += loss.item() * (e - b)
    grads = {k: v.grad for k, v in params.items() if v.grad is not None} if want_grads else {}
    return {'loss': total / x.shape[0]}, grads


def aemsp_forward(sd, hparams, x):
    """AEMSP.forward (aes.py:973-995) -> (x_hat, z, y)."""
    z = encode(sd, hparams, x)
    return decode(sd, hparams, z), z, F.linear(z, sd['projection.weight'])


def aemsp_loss(sd, hparams, x, labels, masks=None, chunk_size=200, want_grads=True):
    """AEMSP.loss (aes.py:997-1077) without the host-side sklearn 'labels_r2' entry:
    mse(x) + msp.alpha * (mse(labels, y_hat) + mse(z, y_hat @ P))."""
    frozen = ('U.weight',)
    params = {k: v.detach().clone().requires_grad_(want_grads and k not in frozen) for k, v in sd.items()}
    vals = {'loss': 0.0, 'loss_mse': 0.0, 'loss_msp': 0.0}
    for b, e in _chunks(x.shape[0], chunk_size):
        x_in = x[b:e]
        x_hat, z, y_hat = aemsp_forward(params, hparams, x_in)
        l_mse = mse(x_in, x_hat, None if masks is None else masks[b:e])
        l_msp = mse(labels[b:e], y_hat) + mse(z, torch.matmul(y_hat, params['projection.weight']))
        loss = l_mse + hparams['msp.alpha'] * l_msp
        if want_grads:
            loss.backward()
        bs = e - b
        vals['loss'] += loss.item() * bs
        vals['loss_mse'] += l_mse.item() * bs
        vals['loss_msp'] += l_msp.item() * bs
    for k in vals:
        vals[k] /= x.shape[0]
    grads = {k: v.grad for k, v in params.items() if v.grad is not None} if want_grads else {}
    return vals, grads


# ------------------------------------------------------------------------------------------------
# PS-VAE (models/vaes.py)
# ------------------------------------------------------------------------------------------------

def psvae_forward(sd, hparams, x, eps=None, use_mean=False):
    """PSVAE.forward (vaes.py:571-601) with the reparameterisation noise injected.

    Returns (x_hat, z, mu, logvar, y_hat).  ``eps`` replaces ``torch.randn_like`` in
    ``reparameterize`` (vaes.py:33-35; std = exp(logvar), reproduced as is).
    """
    n_labels = hparams['n_labels']
    h = encoder_features(sd, hparams, x)
    pre = F.linear(h, sd['encoding.FF.weight'], sd['encoding.FF.bias'])
    y = F.linear(pre, sd['encoding.A.weight'])
    w = F.linear(pre, sd['encoding.B.weight'])
    logvar = F.linear(h, sd['encoding.logvar.weight'], sd['encoding.logvar.bias'])
    mu = torch.cat([y, w], 1)
    z = mu if use_mean else eps * torch.exp(logvar) + mu
    x_hat = decode(sd, hparams, z)
    y_hat = y * sd['encoding.D.weight'] + sd['encoding.D.bias']      # DiagLinear, base.py:70-103
    assert y_hat.shape[1] == n_labels
    return x_hat, z, mu, logvar, y_hat


def psvae_loss(sd, hparams, x, labels, eps, masks=None, labels_masks=None, alpha=None, beta=None,
               kl_anneal=1.0, chunk_size=200, want_grads=True):
    """PSVAE.loss (vaes.py:603-729) for fixed (alpha, beta, kl-anneal) weights.

    Returns (loss dict, grads).  'label_r2' is omitted (sklearn, CPU-side in the reference too,
    vaes.py:711-718); 'loss_data_mse' reproduces the running-sum quirk of vaes.py:705-706.
    """
    alpha = hparams['ps_vae.alpha'] if alpha is None else alpha
    beta = hparams['ps_vae.beta'] if beta is None else beta
    nl = hparams['n_labels']
    frozen = ('encoding.A.weight', 'encoding.B.weight')
    params = {k: v.detach().clone().requires_grad_(want_grads and k not in frozen)
              for k, v in sd.items()}
    keys = ['loss', 'loss_data_ll', 'loss_label_ll', 'loss_zs_kl', 'loss_zu_mi', 'loss_zu_tc',
            'loss_zu_dwkl']
    vals = {k: 0.0 for k in keys}
    vals['loss_data_mse'] = 0.0
    n_pix = int(np.prod(x.shape[1:]))
    for b, e in _chunks(x.shape[0], chunk_size):
        x_in, y_in = x[b:e], labels[b:e]
        x_hat, z, mu, logvar, y_hat = psvae_forward(params, hparams, x_in, eps[b:e])
        t = {}
        t['loss_data_ll'] = gaussian_ll(x_in, x_hat, None if masks is None else masks[b:e])
        t['loss_label_ll'] = gaussian_ll(
            y_in, y_hat, None if labels_masks is None else labels_masks[b:e])
        t['loss_zs_kl'] = kl_div_to_std_normal(mu[:, :nl], logvar[:, :nl])
        t['loss_zu_mi'], t['loss_zu_tc'], t['loss_zu_dwkl'] = decomposed_kl(
            z[:, nl:], mu[:, nl:], logvar[:, nl:])
        t['loss'] = (-t['loss_data_ll'] - alpha * t['loss_label_ll'] + t['loss_zs_kl']
                     + kl_anneal * t['loss_zu_mi'] + beta * t['loss_zu_tc']
                     + kl_anneal * t['loss_zu_dwkl'])
        if want_grads:
            t['loss'].backward()
        bs = e - b
        for k in keys:
            vals[k] += t[k].item() * bs
        vals['loss_data_mse'] += gaussian_ll_to_mse(vals['loss_data_ll'] / bs, n_pix) * bs
    for k in vals:
        vals[k] /= x.shape[0]
    vals['alpha'], vals['beta'] = alpha, beta
    grads = {k: v.grad for k, v in params.items() if v.grad is not None} if want_grads else {}
    return vals, grads


# ------------------------------------------------------------------------------------------------
# multi-session PS-VAE (models/vaes.py:849-1273, 1366-1462; fitting/losses.py:402-513)
# ------------------------------------------------------------------------------------------------

def triplet_loss(z, datasets, margin=1.0):
    """losses.triplet_loss with nn.TripletMarginLoss(margin=1, p=2): written out per session count the way
    the reference enumerates its terms (anchor chunk, positive chunk, (negative session, negative chunk))."""
    tables = {
        2: (3, 3, [(0, 0, 1, 1, 2), (1, 0, 1, 0, 2)]),
        3: (6, 6, [(0, 0, 1, 1, 4), (0, 2, 3, 2, 4), (1, 0, 1, 0, 4), (1, 2, 3, 2, 5), (2, 0, 1, 0, 5),
                   (2, 2, 3, 1, 5)]),
        4: (9, 12, [(0, 0, 1, 1, 6), (0, 2, 3, 2, 6), (0, 4, 5, 3, 6), (1, 0, 1, 0, 6), (1, 2, 3, 2, 7),
                    (1, 4, 5, 3, 7), (2, 0, 1, 0, 7), (2, 2, 3, 1, 7), (2, 4, 5, 3, 8), (3, 0, 1, 0, 8),
                    (3, 2, 3, 1, 8), (3, 4, 5, 2, 8)]),
    }
    ids = np.unique(datasets)
    n_chunks, n_terms, terms = tables[len(ids)]
    shuffled = [np.random.permutation(np.where(datasets == i)[0]) for i in ids]
    m = min(len(s) // n_chunks for s in shuffled)
    ch = [[s[i::n_chunks][:m] for i in range(n_chunks)] for s in shuffled]
    tm = torch.nn.TripletMarginLoss(margin=margin, p=2)
    loss = 0
    for sa, ca, cp, sn, cn in terms:
        loss = loss + tm(z[ch[sa][ca]], z[ch[sa][cp]], z[ch[sn][cn]])
    for sa, ca, cp, _, _ in terms:
        loss = loss + torch.pairwise_distance(z[ch[sa][ca]], z[ch[sa][cp]]).mean()
    return loss / n_terms


def msps_forward(sd, hparams, x, eps=None, use_mean=False):
    """MSPSVAE.forward (vaes.py:893-924) -> (x_hat, z, mu, logvar, y_hat), mu = [z_s, z_b, z_u]."""
    h = encoder_features(sd, hparams, x)
    pre = F.linear(h, sd['encoding.FF.weight'], sd['encoding.FF.bias'])
    z_s = F.linear(pre, sd['encoding.A.weight'])
    z_u = F.linear(pre, sd['encoding.B.weight'])
    z_b = F.linear(pre, sd['encoding.C.weight'], sd['encoding.C.bias'])
    logvar = F.linear(h, sd['encoding.logvar.weight'], sd['encoding.logvar.bias'])
    mu = torch.cat([z_s, z_b, z_u], 1)
    z = mu if use_mean else eps * torch.exp(logvar) + mu
    y_hat = z_s * sd['encoding.D.weight'] + sd['encoding.D.bias']
    return decode(sd, hparams, z), z, mu, logvar, y_hat


def msps_loss(sd, hparams, x, labels, eps, masks=None, sessions=None, want_grads=True):
    """MSPSVAE.loss (vaes.py:926-1077) on the whole batch; ``sessions`` (per-frame session id) adds the triplet
    term of a multi-session batch.  'label_r2' (host-side sklearn) is omitted."""
    nl, nb = hparams['n_labels'], hparams['n_background']
    frozen = ('encoding.A.weight', 'encoding.B.weight', 'encoding.C.weight')
    params = {k: v.detach().clone().requires_grad_(want_grads and k not in frozen) for k, v in sd.items()}
    x_hat, z, mu, logvar, y_hat = msps_forward(params, hparams, x, eps)
    t = {}
    t['loss_data_ll'] = gaussian_ll(x, x_hat, masks)
    t['loss_label_ll'] = gaussian_ll(labels, y_hat)
    t['loss_zs_kl'] = kl_div_to_std_normal(mu[:, :nl], logvar[:, :nl])
    t['loss_zu_mi'], t['loss_zu_tc'], t['loss_zu_dwkl'] = decomposed_kl(
        z[:, nl + nb:], mu[:, nl + nb:], logvar[:, nl + nb:])
    loss = (-t['loss_data_ll'] - hparams['ps_vae.alpha'] * t['loss_label_ll'] + t['loss_zs_kl'] + t['loss_zu_mi']
            + hparams['ps_vae.beta'] * t['loss_zu_tc'] + t['loss_zu_dwkl'])
    if sessions is not None:
        t['loss_triplet'] = triplet_loss(mu[:, nl:nl + nb], sessions)
        loss = loss + hparams['ps_vae.delta'] * t['loss_triplet']
    if want_grads:
        loss.backward()
    vals = {k: v.item() for k, v in t.items()}
    vals['loss'] = loss.item()
    vals['loss_data_mse'] = gaussian_ll_to_mse(vals['loss_data_ll'], int(np.prod(x.shape[1:])))
    grads = {k: v.grad for k, v in params.items() if v.grad is not None} if want_grads else {}
    return vals, grads


# ------------------------------------------------------------------------------------------------
# VAE / beta-TC-VAE (models/vaes.py:38-208, 367-503)
# ------------------------------------------------------------------------------------------------

def vae_forward(sd, hparams, x, eps=None, use_mean=False, dataset=None):
    """VAE.forward (vaes.py:102-129) -> (x_hat, z, mu, logvar); noise injected like psvae_forward."""
    mu, logvar = encode(sd, hparams, x, dataset=dataset)
    z = mu if use_mean else eps * torch.exp(logvar) + mu
    return decode(sd, hparams, z, dataset=dataset), z, mu, logvar


def vae_loss(sd, hparams, x, eps, masks=None, beta=None, chunk_size=200, want_grads=True, dataset=None):
    """VAE.loss (vaes.py:131-208): -ll + beta * KL per chunk."""
    beta = hparams['vae.beta'] if beta is None else beta
    params = {k: v.detach().clone().requires_grad_(want_grads) for k, v in sd.items()}
    vals = {'loss': 0.0, 'loss_ll': 0.0, 'loss_kl': 0.0, 'loss_mse': 0.0}
    n_pix = int(np.prod(x.shape[1:]))
    for b, e in _chunks(x.shape[0], chunk_size):
        x_in = x[b:e]
        x_hat, _, mu, logvar = vae_forward(params, hparams, x_in, eps[b:e], dataset=dataset)
        ll = gaussian_ll(x_in, x_hat, None if masks is None else masks[b:e])
        kl = kl_div_to_std_normal(mu, logvar)
        loss = -ll + beta * kl
        if want_grads:
            loss.backward()
        bs = e - b
        vals['loss'] += loss.item() * bs
        vals['loss_ll'] += ll.item() * bs
        vals['loss_kl'] += kl.item() * bs
        vals['loss_mse'] += gaussian_ll_to_mse(ll.item(), n_pix) * bs
    for k in vals:
        vals[k] /= x.shape[0]
    vals['beta'] = beta
    grads = {k: v.grad for k, v in params.items() if v.grad is not None} if want_grads else {}
    return vals, grads


def cond_vae_forward(sd, hparams, x, labels, eps=None, use_mean=False):
    """ConditionalVAE.forward (vaes.py:241-280) with conditional_encoder = False (with it the reference's
    encoder only widens its first layer for model_class 'cond-ae', aes.py:129-137)."""
    mu, logvar = encode(sd, hparams, x)
    z = mu if use_mean else eps * torch.exp(logvar) + mu
    return decode(sd, hparams, torch.cat((z, labels), dim=1)), z, mu, logvar


def cond_vae_loss(sd, hparams, x, labels, eps, masks=None, beta=None, chunk_size=200, want_grads=True):
    """ConditionalVAE.loss (vaes.py:282-364)."""
    beta = hparams['vae.beta'] if beta is None else beta
    params = {k: v.detach().clone().requires_grad_(want_grads) for k, v in sd.items()}
    vals = {'loss': 0.0, 'loss_ll': 0.0, 'loss_kl': 0.0, 'loss_mse': 0.0}
    n_pix = int(np.prod(x.shape[1:]))
    for b, e in _chunks(x.shape[0], chunk_size):
        x_in = x[b:e]
        x_hat, _, mu, logvar = cond_vae_forward(params, hparams, x_in, labels[b:e], eps[b:e])
        ll = gaussian_ll(x_in, x_hat, None if masks is None else masks[b:e])
        kl = kl_div_to_std_normal(mu, logvar)
        loss = -ll + beta * kl
        if want_grads:
            loss.backward()
        bs = e - b
        vals['loss'] += loss.item() * bs
        vals['loss_ll'] += ll.item() * bs
        vals['loss_kl'] += kl.item() * bs
        vals['loss_mse'] += gaussian_ll_to_mse(ll.item(), n_pix) * bs
    for k in vals:
        vals[k] /= x.shape[0]
    vals['beta'] = beta
    grads = {k: v.grad for k, v in params.items() if v.grad is not None} if want_grads else {}
    return vals, grads


def btcvae_loss(sd, hparams, x, eps, masks=None, beta=None, kl_anneal=1.0, chunk_size=200, want_grads=True):
    """BetaTCVAE.loss (vaes.py:411-503): -ll + kl * MI + beta * TC + kl * DWKL per chunk;
    'loss_mse' reproduces the running-sum quirk of vaes.py:490-491."""
    beta = hparams['beta_tcvae.beta'] if beta is None else beta
    params = {k: v.detach().clone().requires_grad_(want_grads) for k, v in sd.items()}
    keys = ['loss', 'loss_ll', 'loss_mi', 'loss_tc', 'loss_dwkl']
    vals = {k: 0.0 for k in keys}
    vals['loss_mse'] = 0.0
    n_pix = int(np.prod(x.shape[1:]))
    for b, e in _chunks(x.shape[0], chunk_size):
        x_in = x[b:e]
        x_hat, z, mu, logvar = vae_forward(params, hparams, x_in, eps[b:e])
        t = {}
        t['loss_ll'] = gaussian_ll(x_in, x_hat, None if masks is None else masks[b:e])
        t['loss_mi'], t['loss_tc'], t['loss_dwkl'] = decomposed_kl(z, mu, logvar)
        t['loss'] = -t['loss_ll'] + kl_anneal * t['loss_mi'] + beta * t['loss_tc'] + kl_anneal * t['loss_dwkl']
        if want_grads:
            t['loss'].backward()
        bs = e - b
        for k in keys:
            vals[k] += t[k].item() * bs
        vals['loss_mse'] += gaussian_ll_to_mse(vals['loss_ll'] / bs, n_pix) * bs
    for k in vals:
        vals[k] /= x.shape[0]
    vals['beta'] = beta
    grads = {k: v.grad for k, v in params.items() if v.grad is not None} if want_grads else {}
    return vals, grads


# ------------------------------------------------------------------------------------------------
# synthetic parameters with the reference's shapes/names (torch default inits)
# ------------------------------------------------------------------------------------------------

def make_hparams(n_input_channels, y_pixels, x_pixels, n_ae_latents, model_class='ae',
                 n_labels=0, arch=None, conditional_encoder=False, padding_type=None, n_datasets=0):
    """hparams dict as the reference's grid-search mains assemble it (ae_grid_search.py:20-30).
    ``padding_type='valid'`` selects the unpadded variant of the default arch; ``n_datasets > 0`` the
    per-session input / output layers (``fit_sess_io_layers``)."""
    from behavenet_b200.models.ae_model_architecture_generator import (
        load_default_arch, get_handcrafted_dims)
    hp = load_default_arch() if arch is None else dict(arch)
    if padding_type is not None:
        hp['ae_padding_type'] = padding_type
    hp['ae_batch_norm'] = False
    hp['ae_input_dim'] = [n_input_channels, y_pixels, x_pixels]
    hp['n_input_channels'], hp['y_pixels'], hp['x_pixels'] = n_input_channels, y_pixels, x_pixels
    hp['n_ae_latents'] = n_ae_latents
    hp = get_handcrafted_dims(hp, symmetric=True)
    hp['model_class'] = model_class
    hp['model_type'] = 'conv'
    hp['fit_sess_io_layers'] = n_datasets > 0
    if n_datasets > 0:
        hp['n_datasets'] = n_datasets
    if model_class == 'vae':
        hp.update({'vae.beta': 2.0, 'vae.beta_anneal_epochs': 0, 'max_n_epochs': 10, 'variational': True})
    if model_class == 'beta-tcvae':
        # BetaTCVAE.__init__ runs VAE.__init__, which reads vae.beta as well (vaes.py:100, 389)
        hp.update({'beta_tcvae.beta': 5.0, 'beta_tcvae.beta_anneal_epochs': 0, 'vae.beta': 1.0,
                   'vae.beta_anneal_epochs': 0, 'max_n_epochs': 10, 'variational': True})
    if model_class == 'ps-vae':
        hp.update({'n_labels': n_labels, 'ps_vae.alpha': 1000, 'ps_vae.beta': 10,
                   'ps_vae.anneal_epochs': 0, 'max_n_epochs': 10, 'variational': True})
    if model_class == 'msps-vae':
        hp.update({'n_labels': n_labels, 'n_background': 2, 'n_sessions_per_batch': 2, 'ps_vae.alpha': 1000,
                   'ps_vae.beta': 10, 'ps_vae.delta': 50, 'ps_vae.anneal_epochs': 0, 'max_n_epochs': 10,
                   'variational': True})
    if model_class == 'cond-ae':
        hp.update({'n_labels': n_labels, 'conditional_encoder': conditional_encoder})
    if model_class == 'cond-vae':
        hp.update({'n_labels': n_labels, 'conditional_encoder': False, 'vae.beta': 2.0, 'vae.beta_anneal_epochs': 0,
                   'max_n_epochs': 10, 'variational': True})
    if model_class == 'cond-ae-msp':
        hp.update({'n_labels': n_labels, 'msp.alpha': 0.01})
    return hp


def init_state_dict(hparams, seed=0, dtype=torch.float32):
    """Random parameters with the reference's names/shapes, drawn
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's default Conv2d/ConvTranspose2d/Linear
    initialisers, from a private generator under ``seed``."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def uniform(shape, bound):
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    def put(name, w_shape, fan_in, n_bias):
        sd[name + '.weight'] = uniform(w_shape, 1 / math.sqrt(fan_in))
        sd[name + '.bias'] = uniform((n_bias,), 1 / math.sqrt(fan_in))

    c_in = hparams['ae_input_dim'][0]
    cond = hparams.get('model_class') in ('cond-ae', 'cond-vae')
    if hparams.get('model_class') == 'cond-ae' and hparams.get('conditional_encoder', False):
        c_in += hparams['n_labels'] // 2                 # one-hot label images join the frames (aes.py:129-137)
    sess_io = hparams.get('fit_sess_io_layers', False)
    for i, c in enumerate(hparams['ae_encoding_n_channels']):
        k = hparams['ae_encoding_kernel_size'][i]
        if sess_io and i == 0:      # one input layer per session (aes.py:69-80)
            for d in range(hparams['n_datasets']):
                put('encoding.encoder.conv0_sess_io_layers.%i' % d, (c, c_in, k, k), c_in * k * k, c)
        else:
            put('encoding.encoder.conv%i' % i, (c, c_in, k, k), c_in * k * k, c)
        c_in = c
    feat = c_in * hparams['ae_encoding_y_dim'][-1] * hparams['ae_encoding_x_dim'][-1]
    L = hparams['n_ae_latents']
    put('encoding.FF', (L, feat), feat, L)
    if hparams.get('variational', False):
        put('encoding.logvar', (L, feat), feat, L)
    if hparams.get('model_class') == 'ps-vae':
        nl = hparams['n_labels']
        q, _ = torch.linalg.qr(torch.randn(L, L, generator=g, dtype=torch.float64))
        sd['encoding.A.weight'] = q[:nl].to(dtype).contiguous()
        sd['encoding.B.weight'] = q[nl:].to(dtype).contiguous()
        sd['encoding.D.weight'] = uniform((nl,), 1 / math.sqrt(nl))
        sd['encoding.D.bias'] = uniform((nl,), 1 / math.sqrt(nl))
    if hparams.get('model_class') == 'msps-vae':
        nl, nb = hparams['n_labels'], hparams['n_background']
        q, _ = torch.linalg.qr(torch.randn(L, L, generator=g, dtype=torch.float64))
        sd['encoding.A.weight'] = q[:nl].to(dtype).contiguous()
        sd['encoding.B.weight'] = q[nl + nb:].to(dtype).contiguous()
        sd['encoding.C.weight'] = q[nl:nl + nb].to(dtype).contiguous()
        sd['encoding.C.bias'] = uniform((nb,), 1 / math.sqrt(L))
        sd['encoding.D.weight'] = uniform((nl,), 1 / math.sqrt(nl))
        sd['encoding.D.bias'] = uniform((nl,), 1 / math.sqrt(nl))
    if hparams.get('model_class') == 'cond-ae-msp':
        nl = hparams['n_labels']
        sd['projection.weight'] = uniform((nl, L), 1 / math.sqrt(L))
        sd['U.weight'] = uniform((L, L), 1 / math.sqrt(L))
    c0, h0, w0 = hparams['ae_decoding_starting_dim']
    Ld = L + hparams['n_labels'] if cond else L           # hidden_layer_size (aes.py:803)
    put('decoding.FF', (c0 * h0 * w0, Ld), Ld, c0 * h0 * w0)
    c_in = c0
    for i, c in enumerate(hparams['ae_decoding_n_channels']):
        k = hparams['ae_decoding_kernel_size'][i]
        # torch computes ConvTranspose2d fan_in from weight.size(1) = out channels
        if sess_io and i == len(hparams['ae_decoding_n_channels']) - 1:     # aes.py:298-312
            for d in range(hparams['n_datasets']):
                put('decoding.decoder.convtranspose%i_sess_io_layers.%i' % (i, d), (c_in, c, k, k), c * k * k, c)
        else:
            put('decoding.decoder.convtranspose%i' % i, (c_in, c, k, k), c * k * k, c)
        c_in = c
    return sd
