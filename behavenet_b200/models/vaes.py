"""Partitioned-subspace VAE with the reference's object protocol on sm_100a kernels.

Drop-in for ``behavenet.models.vaes.{reparameterize, ConvAEPSEncoder, PSVAE}`` (reference
``behavenet/models/vaes.py:17-35, 1276-1363, 506-846``).  The conv stacks run on the same kernels
as the AE; the latent block (orthogonal A/B projections, diagonal label head D, reparameterisation,
label log-likelihood, KL and decomposed-KL terms and all of their gradients) is one fused C call
per reference chunk (``bn_psvae_latent``), and the pixel log-likelihood is fused into the last
decoder layer's epilogue.
"""

import numpy as np
import torch
from torch import nn

from behavenet_b200 import _lib, parallel
from behavenet_b200.models.aes import AE, ConvAEDecoder, ConvAEEncoder
from behavenet_b200.models.base import DiagLinear
from behavenet_b200.models._engine import CaeDriver, Runtime
from behavenet_b200.fitting.losses import (gaussian_ll, gaussian_ll_to_mse, kl_div_to_std_normal, decomposed_kl,  # noqa: F401
                                           triplet_loss)

__all__ = ['reparameterize', 'VAE', 'BetaTCVAE', 'PSVAE', 'ConvAEPSEncoder']

LN2PI = float(np.log(2 * np.pi))


def reparameterize(mu, logvar, eps=None):
    """Sample eps * exp(logvar) + mu (reference vaes.py:17-35; note std = exp(logvar)).

    ``eps`` (optional) injects the noise, for bit-reproducible comparisons with the reference.
    """
    if eps is None:
        eps = torch.randn_like(mu)
    return eps * torch.exp(logvar) + mu



def _latent_rows(beg, end, n_total, chunks, pre, logvar, eps_s, call, dp):
    """Run the per-chunk latent block for every reference chunk that holds frames of this rank's
    contiguous range [beg, end) and return this rank's rows of (mu, z, gmu_part, glogvar_part, gz_part).

    The MI / TC / DWKL estimators are pairwise over a whole chunk (reference losses.py:321-341), so a rank
    whose frames end inside a chunk needs the other ranks' (FF output, logvar, eps) rows of that chunk:
    with ``dp`` the (n_total, 3 L) table is assembled by one all-reduce(SUM) of zero-padded rows (an
    all-gather that also works on ragged shards and on the gloo backend; 3 L floats per frame).  Every
    rank that shares a chunk evaluates the (tiny) block on all of its rows and keeps its own; the loss
    sums, label predictions and D gradients of a chunk are contributed by its owner only -- the rank
    holding the chunk's first frame -- so the later all-reduce counts them once.
    ``call(c, b, e, owner, pre, logvar, eps, mu, z, gmu, glv, gz)`` evaluates chunk c = frames [b, e) on
    full-batch-indexed tensors."""
    L = pre.shape[1]
    dev = pre.device
    if dp:
        full = torch.zeros(n_total, 3 * L, dtype=torch.float32, device=dev)
        full[beg:end, :L] = pre
        full[beg:end, L:2 * L] = logvar
        full[beg:end, 2 * L:] = eps_s
        parallel.all_reduce_sum(full)
        pre_f, lv_f, eps_f = (full[:, i * L:(i + 1) * L].contiguous() for i in range(3))
    else:
        assert beg == 0 and end == n_total
        pre_f, lv_f, eps_f = pre, logvar, eps_s
    out = [torch.zeros(n_total, L, dtype=torch.float32, device=dev) for _ in range(5)]
    for c, (b, e) in enumerate(chunks):
        if e <= beg or b >= end:
            continue
        call(c, b, e, beg <= b < end, pre_f, lv_f, eps_f, *out)
    return [o[beg:end] for o in out]


class _LatentFn(torch.autograd.Function):
    """Differentiable PS-VAE latent block outside ``loss`` (forward / plotting helpers):
    (pre, logvar, A, B, Dw, Db, eps) -> (mu, z, y_hat).  Backward chains through the same C call
    that ``PSVAE.loss`` uses, with all loss weights zero."""

    @staticmethod
    def forward(ctx, pre, logvar, A, B, Dw, Db, eps):
        n, L = pre.shape
        nl = A.shape[0]
        dev = pre.device
        mu = torch.empty(n, L, device=dev)
        z = torch.empty(n, L, device=dev)
        yhat = torch.empty(n, nl, device=dev)
        terms = torch.zeros(5, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().bn_psvae_latent(
            n, L, nl, pre.data_ptr(), logvar.data_ptr(), A.data_ptr(), _lib.ptr(B), Dw.data_ptr(),
            Db.data_ptr(), _lib.ptr(eps), None, None, 0.0, 0.0, 0.0, 1.0, None, mu.data_ptr(),
            z.data_ptr(), yhat.data_ptr(), terms.data_ptr(), None, None, None, None, None,
            _lib.stream_ptr()), 'bn_psvae_latent')
        ctx.save_for_backward(logvar, A, B, Dw, eps, mu)
        return mu, z, yhat

    @staticmethod
    def backward(ctx, gmu, gz, gy):
        logvar, A, B, Dw, eps, mu = ctx.saved_tensors
        n, L = logvar.shape
        nl = A.shape[0]
        dev = logvar.device
        gmu_part = torch.zeros(n, L, device=dev) if gmu is None else gmu.contiguous().clone()
        gDw = gDb = None
        if gy is not None:
            gmu_part[:, :nl] += gy * Dw
            gDw = (gy * mu[:, :nl]).sum(0)
            gDb = gy.sum(0)
        gz_dec = None if gz is None else gz.contiguous()
        gpre = torch.empty(n, L, device=dev)
        glv = torch.empty(n, L, device=dev)
        _lib.check(_lib.lib().bn_psvae_latent_bwd(
            n, L, nl, A.data_ptr(), _lib.ptr(B), _lib.ptr(eps), logvar.data_ptr(), _lib.ptr(gz_dec),
            gmu_part.data_ptr(), None, None, gpre.data_ptr(), glv.data_ptr(), _lib.stream_ptr()),
            'bn_psvae_latent_bwd')
        return gpre, glv, None, None, gDw, gDb, None


class ConvAEPSEncoder(ConvAEEncoder):
    """Encoder that separates the label-related subspace (reference vaes.py:1276-1363)."""

    def __init__(self, hparams):
        super().__init__(hparams)
        n_latents = self.hparams['n_ae_latents']
        n_labels = self.hparams['n_labels']
        self.A = nn.Linear(n_latents, n_labels, bias=False)
        self.B = nn.Linear(n_latents, n_latents - n_labels, bias=False)
        self.D = DiagLinear(n_labels, bias=True)
        # frozen orthogonal projections, drawn from numpy's global RNG like the reference
        # (vaes.py:1294-1302)
        from scipy.stats import ortho_group
        m = ortho_group.rvs(dim=n_latents).astype('float32')
        with torch.no_grad():
            self.A.weight = nn.Parameter(torch.from_numpy(m[:n_labels, :]), requires_grad=False)
            self.B.weight = nn.Parameter(torch.from_numpy(m[n_labels:, :]), requires_grad=False)

    def __str__(self):
        s = 'Encoder architecture:\n'
        i = -1
        for i, module in enumerate(self.encoder):
            s += '    {:02d}: {}\n'.format(i, module)
        s += '    {:02d}: {}\n'.format(i + 1, self.FF)
        s += '    {:02d}: {} (to constrained latents)\n'.format(i + 1, self.A)
        s += '    {:02d}: {} (to unconstrained latents)\n'.format(i + 1, self.B)
        s += '    {:02d}: {} (constrained latents to labels)\n'.format(i + 1, self.D)
        return s

    def forward(self, x, dataset=None):
        """(y, w, logvar, pool_idx, output_size) (reference vaes.py:1319-1363): ``logvar`` comes
        from the flattened conv features, y / w are the A / B projections of the FF output."""
        pre, logvar = self._heads(x, dataset)
        n_labels = self.hparams['n_labels']
        mu, _, _ = _LatentFn.apply(pre, logvar, self.A.weight, self.B.weight, self.D.weight,
                                   self.D.bias, None)
        return mu[:, :n_labels], mu[:, n_labels:], logvar, [], []


class VAE(AE):
    """Variational autoencoder / beta-VAE (reference vaes.py:38-208): ``-ll + beta * KL``.

    The conv stacks run on the AE kernels; the latent block is the PS-VAE C call with every latent
    in the "supervised" slice (A = I, no labels), whose analytic KL weight carries ``vae.beta``."""

    def __init__(self, hparams):
        if hparams['model_type'] == 'linear':
            raise NotImplementedError
        hparams['variational'] = True
        super().__init__(hparams)
        anneal_epochs = self.hparams.get('vae.beta_anneal_epochs', 0)
        self.curr_epoch = 0  # must be modified by training script
        tail = np.ones(hparams['max_n_epochs'] + 1)
        if anneal_epochs > 0:
            self.beta_vals = np.append(np.linspace(0, hparams['vae.beta'], anneal_epochs), tail)
        else:
            self.beta_vals = hparams['vae.beta'] * tail

    def forward(self, x, dataset=None, use_mean=False, eps=None, **kwargs):
        """(x_hat, z, mu, logvar) (reference vaes.py:102-129); ``eps`` optionally injects the noise."""
        mu, logvar, pool_idx, outsize = self.encoding(x, dataset=dataset)
        z = mu if use_mean else reparameterize(mu, logvar, None if eps is None else eps.to(mu.device))
        x_hat = self.decoding(z, pool_idx, outsize, dataset=dataset)
        return x_hat, z, mu, logvar

    # -- latent-block configuration: (n "supervised" dims, alpha, beta_tc, kl_w, kl_s_w) ----------
    def _latent_weights(self):
        L = self.hparams['n_ae_latents']
        return L, 0.0, 0.0, 0.0, float(self.beta_vals[self.curr_epoch])

    def _identity(self, device):
        L = self.hparams['n_ae_latents']
        buf = self._rt.bufs.get('eye')
        if buf is None or buf[0].device != device:
            buf = (torch.eye(L, device=device), torch.ones(L, device=device), torch.zeros(L, device=device))
            self._rt.bufs['eye'] = buf
        return buf

    def _elbo_pass(self, data, accumulate_grad, chunk_size, eps, dataset=None):
        """Shared fused pass of the VAE family: per reference chunk the pixel sum of squares and the
        latent terms [unused, analytic KL, MI, TC, DWKL] (sums over the chunk's frames), gradients
        accumulated into ``.grad``.  Returns (host array (n_chunks, 6), chunks)."""
        x = data['images'][0]
        m = data['masks'][0] if 'masks' in data else None
        drv, rt = self._driver, self._rt
        x = drv._check_input(x, "data['images'][0]", drv.img)
        if m is not None:
            m = drv._check_input(m.to(torch.float32), "data['masks'][0]", drv.img)
        n_total = x.shape[0]
        chunks = [(b, min(b + chunk_size, n_total)) for b in range(0, n_total, chunk_size)]
        n_chunks = len(chunks)
        L = self.hparams['n_ae_latents']
        nl, alpha, beta, kl, kls = self._latent_weights()
        device = x.device
        if 'shard' in data:
            raise NotImplementedError("data['shard'] (rank-local staging) is only supported by AE.loss; the VAE "
                                      'family shards the full batch itself when model.data_parallel is set')
        dp = self.data_parallel and parallel.enabled()
        beg, end = parallel.shard_range(n_total) if dp else (0, n_total)      # contiguous frames per rank
        n = end - beg
        params = self._kernel_params(dataset)
        sse = torch.zeros(n_chunks, dtype=torch.float64, device=device)
        terms = torch.zeros(n_chunks, 5, dtype=torch.float64, device=device)
        lib = _lib.lib()
        pending = None
        if n > 0:
            xs = x[beg:end]
            ms = None if m is None else m[beg:end]
            if eps is None:
                eps_s = torch.randn(n, L, dtype=torch.float32, device=device)
            else:
                eps_s = eps.to(device=device, dtype=torch.float32)[beg:end].contiguous()
            eye, ones, zeros = self._identity(device)
            packed = drv.packed(rt, params, device)
            ws = drv.workspace(rt, n, device)
            pre, logvar = drv.encode(xs, params, packed, ws, True)
        else:
            pre = logvar = eps_s = torch.zeros(0, L, dtype=torch.float32, device=device)
        if n > 0 or dp:
            eye, ones, zeros = self._identity(device)
            yhat = torch.empty(max(chunk_size * nl, 1), device=device)
            scratch = torch.zeros(5, dtype=torch.float64, device=device)
            lws = torch.empty(lib.bn_psvae_latent_workspace_bytes(chunk_size, L), dtype=torch.uint8, device=device)
            A = eye.data_ptr() if nl > 0 else None
            B = eye.data_ptr() if nl < L else None

            def call(c, b, e, owner, pre_f, lv_f, eps_f, mu_f, z_f, gmu_f, glv_f, gz_f):
                sl = slice(b, e)
                _lib.check(lib.bn_psvae_latent(
                    e - b, L, nl, pre_f[sl].data_ptr(), lv_f[sl].data_ptr(), A, B,
                    ones.data_ptr() if nl > 0 else None, zeros.data_ptr() if nl > 0 else None,
                    eps_f[sl].data_ptr(), None, None, float(alpha), float(beta), float(kl), float(kls),
                    lws.data_ptr(), mu_f[sl].data_ptr(), z_f[sl].data_ptr(), yhat.data_ptr() if nl > 0 else None,
                    (terms[c] if owner else scratch).data_ptr(), gmu_f[sl].data_ptr(), glv_f[sl].data_ptr(),
                    gz_f[sl].data_ptr(), None, None, _lib.stream_ptr()), 'bn_psvae_latent')
            mu, z, gmu_p, glv_p, gz_p = _latent_rows(beg, end, n_total, chunks, pre, logvar, eps_s, call, dp)
        if n > 0:
            grads = self._grad_table(params) if accumulate_grad else None
            drv.decode(z, params, packed, ws, want_xhat=False, target=xs, mask=ms, chunk_size=chunk_size,
                       frame_offset=beg, n_total=n_total, grad_coef=1.0, sse=sse)
            if accumulate_grad:
                gz_dec = drv.decode_bwd(n, None, params, packed, ws, grads, device)
                pending = self._allreduce_begin(params, 'dec') if dp else None
                gpre = torch.empty(n, L, device=device)
                glv = torch.empty(n, L, device=device)
                _lib.check(lib.bn_psvae_latent_bwd(
                    n, L, nl, A, B, eps_s.data_ptr(), logvar.data_ptr(), gz_dec.data_ptr(), gmu_p.data_ptr(),
                    glv_p.data_ptr(), gz_p.data_ptr(), gpre.data_ptr(), glv.data_ptr(), _lib.stream_ptr()),
                    'bn_psvae_latent_bwd')
                drv.encode_bwd(xs, gpre, glv, params, packed, ws, grads)
        elif accumulate_grad:
            self._grad_table(params)
        stats = torch.cat([sse[:, None], terms], 1)
        if dp:
            if accumulate_grad:
                self._allreduce(params, stats, pending)
            else:
                parallel.all_reduce_sum(stats)
        return stats.cpu().numpy(), chunks        # the one device->host read of the call

    def loss(self, data, dataset=0, accumulate_grad=True, chunk_size=200, eps=None):
        """ELBO loss with the reference's chunk semantics (vaes.py:131-208); same dict of floats."""
        host, chunks = self._elbo_pass(data, accumulate_grad, chunk_size, eps, dataset)
        beta = self.beta_vals[self.curr_epoch]
        n_pix = float(np.prod(self._driver.img))
        n_total = chunks[-1][1]
        vals = {'loss': 0.0, 'loss_ll': 0.0, 'loss_kl': 0.0, 'loss_mse': 0.0}
        for c, (b, e) in enumerate(chunks):
            bs = e - b
            ll = -0.5 * LN2PI * n_pix - 0.5 * host[c, 0] / bs
            klv = host[c, 2] / bs
            vals['loss'] += (-ll + beta * klv) * bs
            vals['loss_ll'] += ll * bs
            vals['loss_kl'] += klv * bs
            vals['loss_mse'] += ((ll + 0.5 * LN2PI * n_pix) * -2.0 / n_pix) * bs
        for k in vals:
            vals[k] /= n_total
        vals['beta'] = beta
        return vals


class ConditionalVAE(VAE):
    """Conditional VAE (reference vaes.py:211-364): labels are concatenated to the sampled latents in front
    of the decoder's FF layer.  The conv stacks run through the autograd bridges of the kernels; the Gaussian
    log-likelihood and the analytic KL are torch ops on the outputs, chunk by chunk like the reference."""

    def build_model(self):
        self.hparams['hidden_layer_size'] = self.hparams['n_ae_latents'] + self.hparams['n_labels']
        self.encoding = ConvAEEncoder(self.hparams)
        self.decoding = ConvAEDecoder(self.hparams)
        self._driver = self.encoding._driver
        self._rt = Runtime()

    def invalidate_packed(self):
        self.encoding._rt.packed_key = None
        self.decoding._rt.packed_key = None

    def forward(self, x, dataset=None, labels=None, labels_2d=None, use_mean=False, eps=None, **kwargs):
        """(x_hat, z, mu, logvar) (vaes.py:241-280)."""
        if self.hparams['conditional_encoder']:
            x = torch.cat((x, labels_2d), dim=1)
        mu, logvar, pool_idx, outsize = self.encoding(x, dataset=dataset)
        z = mu if use_mean else reparameterize(mu, logvar, None if eps is None else eps.to(mu.device))
        x_hat = self.decoding(torch.cat((z, labels), dim=1), pool_idx, outsize, dataset=dataset)
        return x_hat, z, mu, logvar

    def loss(self, data, dataset=0, accumulate_grad=True, chunk_size=200, eps=None):
        """{'loss', 'loss_ll', 'loss_kl', 'loss_mse', 'beta'} (vaes.py:282-364)."""
        x, y = data['images'][0], data['labels'][0]
        m = data['masks'][0] if 'masks' in data else None
        y2d = data['labels_sc'][0] if self.hparams['conditional_encoder'] else None
        beta = self.beta_vals[self.curr_epoch]
        n = x.shape[0]
        n_dims = int(np.prod(x.shape[1:]))
        vals = {'loss': 0.0, 'loss_ll': 0.0, 'loss_kl': 0.0, 'loss_mse': 0.0}
        for b in range(0, n, chunk_size):
            e = min(b + chunk_size, n)
            x_hat, _, mu, logvar = self.forward(
                x[b:e], dataset=dataset, labels=y[b:e], labels_2d=None if y2d is None else y2d[b:e],
                eps=None if eps is None else eps[b:e])
            ll = gaussian_ll(x[b:e], x_hat, None if m is None else m[b:e])
            kl = kl_div_to_std_normal(mu, logvar)
            loss = -ll + beta * kl
            if accumulate_grad:
                loss.backward()
            vals['loss'] += loss.item() * (e - b)
            vals['loss_ll'] += ll.item() * (e - b)
            vals['loss_kl'] += kl.item() * (e - b)
            vals['loss_mse'] += float(gaussian_ll_to_mse(ll.item(), n_dims)) * (e - b)
        out = {k: v / n for k, v in vals.items()}
        out['beta'] = beta
        return out


class BetaTCVAE(VAE):
    """beta-TC-VAE (reference vaes.py:367-503): ``-ll + kl * MI + beta * TC + kl * DWKL`` with the
    minibatch estimators of losses.py:284-372 over every latent (B = I, no supervised slice)."""

    def __init__(self, hparams):
        super().__init__(hparams)
        anneal_epochs = self.hparams.get('beta_tcvae.beta_anneal_epochs', 0)
        self.curr_epoch = 0
        beta = hparams['beta_tcvae.beta']
        tail = np.ones(hparams['max_n_epochs'] + 1)
        if anneal_epochs > 0:
            self.beta_vals = np.append(np.linspace(0, beta, anneal_epochs), beta * tail)
            self.kl_anneal_vals = np.append(np.linspace(0, 1, anneal_epochs), tail)
        else:
            self.beta_vals = beta * tail
            self.kl_anneal_vals = tail

    def _latent_weights(self):
        return 0, 0.0, float(self.beta_vals[self.curr_epoch]), float(self.kl_anneal_vals[self.curr_epoch]), 0.0

    def loss(self, data, dataset=0, accumulate_grad=True, chunk_size=200, eps=None):
        host, chunks = self._elbo_pass(data, accumulate_grad, chunk_size, eps, dataset)
        beta = self.beta_vals[self.curr_epoch]
        kl = self.kl_anneal_vals[self.curr_epoch]
        n_pix = float(np.prod(self._driver.img))
        n_total = chunks[-1][1]
        keys = ['loss', 'loss_ll', 'loss_mi', 'loss_tc', 'loss_dwkl']
        vals = {k: 0.0 for k in keys}
        vals['loss_mse'] = 0.0
        for c, (b, e) in enumerate(chunks):
            bs = e - b
            t = {'loss_ll': -0.5 * LN2PI * n_pix - 0.5 * host[c, 0] / bs, 'loss_mi': host[c, 3] / bs,
                 'loss_tc': host[c, 4] / bs, 'loss_dwkl': host[c, 5] / bs}
            t['loss'] = -t['loss_ll'] + kl * t['loss_mi'] + beta * t['loss_tc'] + kl * t['loss_dwkl']
            for k in keys:
                vals[k] += t[k] * bs
            # the reference converts the RUNNING sum (vaes.py:490-491); reproduced as is
            llc = vals['loss_ll'] / bs + 0.5 * LN2PI * n_pix
            vals['loss_mse'] += (llc * -2.0 / n_pix) * bs
        for k in vals:
            vals[k] /= n_total
        vals['beta'] = beta
        return vals


class PSVAE(AE):
    """Partitioned subspace variational autoencoder (reference vaes.py:506-846)."""

    def __init__(self, hparams):
        if hparams['model_type'] == 'linear':
            raise NotImplementedError
        if hparams['n_ae_latents'] < hparams['n_labels']:
            raise ValueError('PS-VAE model must contain at least as many latents as labels')
        self.n_latents = hparams['n_ae_latents']
        self.n_labels = hparams['n_labels']
        hparams['variational'] = True
        super().__init__(hparams)
        # annealing tables indexed by curr_epoch (reference vaes.py:536-553)
        anneal_epochs = self.hparams.get('ps_vae.anneal_epochs', 0)
        self.curr_epoch = 0
        beta = hparams['ps_vae.beta']
        tail = np.ones(hparams['max_n_epochs'] + 1)
        if anneal_epochs > 0:
            self.beta_vals = np.append(np.linspace(0, beta, anneal_epochs), beta * tail)
            self.kl_anneal_vals = np.append(np.linspace(0, 1, anneal_epochs), tail)
        else:
            self.beta_vals = beta * tail
            self.kl_anneal_vals = tail

    def build_model(self):
        self.hparams['hidden_layer_size'] = self.hparams['n_ae_latents']
        if self.model_type == 'conv':
            self.encoding = ConvAEPSEncoder(self.hparams)
            self.decoding = ConvAEDecoder(self.hparams)
        elif self.model_type == 'linear':
            raise NotImplementedError
        else:
            raise ValueError('"%s" is an invalid model_type' % self.model_type)
        from behavenet_b200.models._engine import Runtime
        self._driver = self.encoding._driver
        self._rt = Runtime()

    def _extra_trainable(self):
        return [self.encoding.D.weight, self.encoding.D.bias]

    def forward(self, x, dataset=None, use_mean=False, eps=None, **kwargs):
        """(x_hat, z, mu, logvar, y_hat) (reference vaes.py:571-601).  ``eps`` optionally injects
        the reparameterisation noise."""
        enc = self.encoding
        pre, logvar = enc._heads(x)
        if use_mean:
            noise = None
        else:
            noise = torch.randn_like(pre) if eps is None else eps.to(pre.device).contiguous()
        mu, z, y_hat = _LatentFn.apply(pre, logvar, enc.A.weight, enc.B.weight, enc.D.weight,
                                       enc.D.bias, noise)
        x_hat = self.decoding(z, [], [], dataset=dataset)
        return x_hat, z, mu, logvar, y_hat

    def loss(self, data, dataset=0, accumulate_grad=True, chunk_size=200, eps=None):
        """Decomposed-ELBO loss of the PS-VAE with the reference's chunk semantics
        (reference vaes.py:603-729).  Returns the same dict of python floats.

        Conv stacks run once over the whole (rank-local) batch; the latent block runs once per
        reference chunk because the MI / TC / DWKL estimators are pairwise over the chunk
        (losses.py:321-351).  ``eps`` (n, n_latents) optionally injects the sampling noise.
        """
        x = data['images'][0]
        y = data['labels'][0]
        m = data['masks'][0] if 'masks' in data else None
        nm = data['labels_masks'][0] if 'labels_masks' in data else None
        drv, rt, enc = self._driver, self._rt, self.encoding
        x = drv._check_input(x, "data['images'][0]", drv.img)
        y = drv._check_input(y.to(torch.float32), "data['labels'][0]", (self.n_labels,))
        if m is not None:
            m = drv._check_input(m.to(torch.float32), "data['masks'][0]", drv.img)
        if nm is not None:
            nm = drv._check_input(nm.to(torch.float32), "data['labels_masks'][0]", (self.n_labels,))
        n_total = x.shape[0]
        chunks = [(b, min(b + chunk_size, n_total)) for b in range(0, n_total, chunk_size)]
        n_chunks = len(chunks)
        L, nl = self.n_latents, self.n_labels
        alpha = self.hparams['ps_vae.alpha']
        beta = self.beta_vals[self.curr_epoch]
        kl = self.kl_anneal_vals[self.curr_epoch]
        device = x.device
        if 'shard' in data:
            raise NotImplementedError("data['shard'] (rank-local staging) is only supported by AE.loss; PSVAE "
                                      'shards the full batch itself when model.data_parallel is set')
        dp = self.data_parallel and parallel.enabled()
        # contiguous frames per rank, balanced by frames; chunks that span ranks exchange their
        # (FF output, logvar, eps) rows (reference losses.py:321-341)
        beg, end = parallel.shard_range(n_total) if dp else (0, n_total)
        n = end - beg
        params = self._kernel_params(dataset)
        # per chunk: [sse_pixels] ; [label sumsq, zs_kl, mi, tc, dwkl]
        sse = torch.zeros(n_chunks, dtype=torch.float64, device=device)
        terms = torch.zeros(n_chunks, 5, dtype=torch.float64, device=device)
        y_hat_all = torch.zeros(n_total, nl, dtype=torch.float32, device=device)
        lib = _lib.lib()
        pending = None
        if n > 0:
            xs = x[beg:end]
            ms = None if m is None else m[beg:end]
            if eps is None:
                eps_s = torch.randn(n, L, dtype=torch.float32, device=device)
            else:
                eps_s = eps.to(device=device, dtype=torch.float32)[beg:end].contiguous()
            packed = drv.packed(rt, params, device)
            ws = drv.workspace(rt, n, device)
            pre, logvar = drv.encode(xs, params, packed, ws, True)
        else:
            pre = logvar = eps_s = torch.zeros(0, L, dtype=torch.float32, device=device)
        grads = self._grad_table(params) if accumulate_grad else None
        if n > 0 or dp:
            D = enc.D
            scratch = torch.zeros(5, dtype=torch.float64, device=device)
            yhat_scratch = torch.empty(chunk_size, nl, dtype=torch.float32, device=device)
            lws = torch.empty(lib.bn_psvae_latent_workspace_bytes(chunk_size, L), dtype=torch.uint8,
                              device=device)

            def call(c, b, e, owner, pre_f, lv_f, eps_f, mu_f, z_f, gmu_f, glv_f, gz_f):
                sl = slice(b, e)
                own_grad = owner and accumulate_grad
                _lib.check(lib.bn_psvae_latent(
                    e - b, L, nl, pre_f[sl].data_ptr(), lv_f[sl].data_ptr(), enc.A.weight.data_ptr(),
                    _lib.ptr(enc.B.weight) if L > nl else None, D.weight.data_ptr(), D.bias.data_ptr(),
                    eps_f[sl].data_ptr(), y[sl].data_ptr(), None if nm is None else nm[sl].data_ptr(),
                    float(alpha), float(beta), float(kl), 1.0, lws.data_ptr(), mu_f[sl].data_ptr(),
                    z_f[sl].data_ptr(), (y_hat_all[sl] if owner else yhat_scratch).data_ptr(),
                    (terms[c] if owner else scratch).data_ptr(),
                    gmu_f[sl].data_ptr(), glv_f[sl].data_ptr(), gz_f[sl].data_ptr(),
                    D.weight.grad.data_ptr() if own_grad and D.weight.requires_grad else None,
                    D.bias.grad.data_ptr() if own_grad and D.bias.requires_grad else None,
                    _lib.stream_ptr()), 'bn_psvae_latent')
            mu, z, gmu_p, glv_p, gz_p = _latent_rows(beg, end, n_total, chunks, pre, logvar, eps_s, call, dp)
        if n > 0:
            # pixel log-likelihood fused in the decoder epilogue: d(-ll)/dxhat = (xhat - x) m / len
            drv.decode(z, params, packed, ws, want_xhat=False, target=xs, mask=ms,
                       chunk_size=chunk_size, frame_offset=beg, n_total=n_total, grad_coef=1.0,
                       sse=sse)
        stats_work = None
        if dp and parallel.overlap_enabled():
            # every loss term and the label predictions are final after the forward pass: their all-reduces run under
            # the backward pass (same order on every rank, also on one that owns no frames)
            stats = torch.cat([sse[:, None], terms], 1)
            stats_work = (parallel.all_reduce_sum_async(stats), parallel.all_reduce_sum_async(y_hat_all))
        if n > 0:
            if accumulate_grad:
                gz_dec = drv.decode_bwd(n, None, params, packed, ws, grads, device)
                pending = self._allreduce_begin(params, 'dec') if dp else None
                gpre = torch.empty(n, L, device=device)
                glv = torch.empty(n, L, device=device)
                _lib.check(lib.bn_psvae_latent_bwd(
                    n, L, nl, enc.A.weight.data_ptr(), _lib.ptr(enc.B.weight) if L > nl else None,
                    eps_s.data_ptr(), logvar.data_ptr(), gz_dec.data_ptr(), gmu_p.data_ptr(),
                    glv_p.data_ptr(), gz_p.data_ptr(), gpre.data_ptr(), glv.data_ptr(),
                    _lib.stream_ptr()), 'bn_psvae_latent_bwd')
                drv.encode_bwd(xs, gpre, glv, params, packed, ws, grads)
        if dp:
            if n == 0 and accumulate_grad:
                pending = self._allreduce_begin(params, 'dec')
            if stats_work is None:
                stats = torch.cat([sse[:, None], terms], 1)
            if accumulate_grad:
                self._allreduce(params, stats if stats_work is None else None, pending)
            elif stats_work is None:
                parallel.all_reduce_sum(stats)
            if stats_work is None:
                parallel.all_reduce_sum(y_hat_all)
            else:
                stats_work[0].wait()
                stats_work[1].wait()
            sse, terms = stats[:, 0], stats[:, 1:]
        # ---- one device->host read, then the reference's bookkeeping (vaes.py:700-729)
        host = torch.cat([sse[:, None], terms], 1).cpu().numpy()
        n_pix = float(np.prod(drv.img))
        keys = ['loss', 'loss_data_ll', 'loss_label_ll', 'loss_zs_kl', 'loss_zu_mi', 'loss_zu_tc',
                'loss_zu_dwkl']
        vals = {k: 0.0 for k in keys}
        vals['loss_data_mse'] = 0.0
        for c, (b, e) in enumerate(chunks):
            bs = e - b
            t = {}
            t['loss_data_ll'] = -0.5 * LN2PI * n_pix - 0.5 * host[c, 0] / bs
            t['loss_label_ll'] = -0.5 * LN2PI * nl - 0.5 * host[c, 1] / bs
            t['loss_zs_kl'] = host[c, 2] / bs
            t['loss_zu_mi'] = host[c, 3] / bs
            t['loss_zu_tc'] = host[c, 4] / bs
            t['loss_zu_dwkl'] = host[c, 5] / bs
            t['loss'] = (-t['loss_data_ll'] - alpha * t['loss_label_ll'] + t['loss_zs_kl']
                         + kl * t['loss_zu_mi'] + beta * t['loss_zu_tc'] + kl * t['loss_zu_dwkl'])
            for k in keys:
                vals[k] += t[k] * bs
            # the reference converts the RUNNING sum (vaes.py:705-706); reproduced as is
            llc = vals['loss_data_ll'] / bs + 0.5 * LN2PI * n_pix
            vals['loss_data_mse'] += (llc * -2.0 / n_pix) * bs
        for k in vals:
            vals[k] /= n_total
        vals['alpha'] = alpha
        vals['beta'] = beta
        # variance-weighted R^2 on the host, as the reference does with sklearn (vaes.py:709-718)
        from ..fitting.losses import r2_variance_weighted
        y_np = y.cpu().numpy()
        yh_np = y_hat_all.cpu().numpy()
        if nm is not None:
            n_np = nm.cpu().numpy()
            vals['label_r2'] = r2_variance_weighted(y_np[n_np == 1], yh_np[n_np == 1])
        else:
            vals['label_r2'] = r2_variance_weighted(y_np, yh_np)
        return vals

    # -- helpers used by the reference's plotting code (vaes.py:731-846); tiny tensors ----------
    def get_predicted_labels(self, x, dataset=None, use_mean=True):
        y, w, logvar, _, _ = self.encoding(x, dataset=dataset)
        if not use_mean:
            y = reparameterize(y, logvar[:, :self.n_labels])
        return self.encoding.D(y)

    def get_transformed_latents(self, inputs, dataset=None, as_numpy=True):
        if not isinstance(inputs, torch.Tensor):
            inputs = torch.Tensor(inputs)
        dev = self.encoding.D.weight.device
        inputs = inputs.to(dev)
        if inputs.dim() == 2:
            y_og, w_og = inputs[:, :self.n_labels], inputs[:, self.n_labels:]
        elif inputs.dim() == 4:
            y_og, w_og, _, _, _ = self.encoding(inputs, dataset=dataset)
        else:
            raise ValueError('"inputs" must be 2 or 4-dimensional tensor')
        y_new = self.encoding.D(y_og)
        latents = torch.cat([y_new, w_og], axis=1)
        return latents.cpu().detach().numpy() if as_numpy else latents

    def get_inverse_transformed_latents(self, inputs, dataset=None, as_numpy=True):
        if not isinstance(inputs, torch.Tensor):
            inputs = torch.Tensor(inputs)
        dev = self.encoding.D.weight.device
        inputs = inputs.to(dev)
        if inputs.dim() == 2:
            y_og, w_og = inputs[:, :self.n_labels], inputs[:, self.n_labels:]
        elif inputs.dim() == 4:
            y_og, w_og, _, _, _ = self.encoding(inputs, dataset=dataset)
        else:
            raise ValueError('"inputs" must be 2 or 4-dimensional tensor')
        y_new = (y_og - self.encoding.D.bias) / self.encoding.D.weight
        latents = torch.cat([y_new, w_og], axis=1)
        return latents.cpu().detach().numpy() if as_numpy else latents


# ------------------------------------------------------------------------------------------------
# multi-session PS-VAE (reference vaes.py:849-1462)
# ------------------------------------------------------------------------------------------------

class ConvAEMSPSEncoder(ConvAEEncoder):
    """Encoder with supervised (A), unsupervised (B) and session-background (C) subspaces (reference
    vaes.py:1366-1462): rows of one fixed random orthogonal matrix; C carries a trainable bias."""

    def __init__(self, hparams):
        super().__init__(hparams)
        L, nl, nb = self.hparams['n_ae_latents'], self.hparams['n_labels'], self.hparams['n_background']
        self.A = nn.Linear(L, nl, bias=False)
        self.B = nn.Linear(L, L - nl - nb, bias=False)
        self.C = nn.Linear(L, nb, bias=True)
        self.D = DiagLinear(nl, bias=True)
        from scipy.stats import ortho_group
        m = ortho_group.rvs(dim=L).astype('float32')
        with torch.no_grad():
            self.A.weight = nn.Parameter(torch.from_numpy(m[:nl, :]), requires_grad=False)
            self.B.weight = nn.Parameter(torch.from_numpy(m[nl + nb:, :]), requires_grad=False)
            self.C.weight = nn.Parameter(torch.from_numpy(m[nl:nl + nb, :]), requires_grad=False)

    def __str__(self):
        s = 'Encoder architecture:\n'
        i = -1
        for i, module in enumerate(self.encoder):
            s += '    {:02d}: {}\n'.format(i, module)
        s += '    {:02d}: {}\n'.format(i + 1, self.FF)
        s += '    {:02d}: {} (to supervised latents)\n'.format(i + 1, self.A)
        s += '    {:02d}: {} (to unsupervised latents)\n'.format(i + 1, self.B)
        s += '    {:02d}: {} (to background latents)\n'.format(i + 1, self.C)
        s += '    {:02d}: {} (supervised latents to labels)\n'.format(i + 1, self.D)
        return s

    def forward(self, x, dataset=None):
        """(z_s, z_b, z, logvar, pool_idx, output_size); the conv stack and both heads run in the kernels,
        the three small projections are torch ops on (n, latents) tensors."""
        pre, logvar = self._heads(x, dataset)
        return self.A(pre), self.C(pre), self.B(pre), logvar, [], []


class MSPSVAE(PSVAE):
    """Multi-session PS-VAE (reference vaes.py:849-1273): PS-VAE terms on the whole batch plus a triplet
    loss that pushes the background latents of different sessions apart."""

    def __init__(self, hparams):
        if hparams['n_sessions_per_batch'] == 1:
            raise ValueError('must choose "n_sessions_per_batch" > 1 in hparams')
        hparams['n_background'] = hparams.get('n_background', 4)
        super().__init__(hparams)
        self.TripletLoss = nn.TripletMarginLoss(margin=1.0, p=2)

    def build_model(self):
        self.hparams['hidden_layer_size'] = self.hparams['n_ae_latents']
        if self.model_type != 'conv':
            raise NotImplementedError
        self.encoding = ConvAEMSPSEncoder(self.hparams)
        self.decoding = ConvAEDecoder(self.hparams)
        self._driver = self.encoding._driver
        self._rt = Runtime()

    def invalidate_packed(self):
        self.encoding._rt.packed_key = None
        self.decoding._rt.packed_key = None

    def forward(self, x, dataset=None, use_mean=False, eps=None, **kwargs):
        """(x_hat, z, mu, logvar, y_hat) with mu = [z_s, z_b, z_u] (vaes.py:893-924)."""
        z_s, z_b, z_u, logvar, pool_idx, outsize = self.encoding(x, dataset=dataset)
        mu = torch.cat([z_s, z_b, z_u], dim=1)
        z = mu if use_mean else reparameterize(mu, logvar, None if eps is None else eps.to(mu.device))
        x_hat = self.decoding(z, pool_idx, outsize, dataset=dataset)
        return x_hat, z, mu, logvar, self.encoding.D(z_s)

    def loss(self, datas, dataset=None, accumulate_grad=True, chunk_size=None, eps=None):
        """One pass over the whole batch (a dict, or a list of per-session dicts with ``dataset`` the list
        of their session ids, which adds the triplet term) (vaes.py:926-1077)."""
        from ..fitting.losses import r2_variance_weighted
        multi = isinstance(datas, list)
        if multi:
            def cat(key):
                return torch.cat([d[key][0] for d in datas], dim=0) if key in datas[0] else None
            x, y, m, n = cat('images'), cat('labels'), cat('masks'), cat('labels_masks')
            sessions = np.concatenate([s * np.ones(datas[i]['images'].shape[1]) for i, s in enumerate(dataset)])
        else:
            x, y = datas['images'][0], datas['labels'][0]
            m = datas['masks'][0] if 'masks' in datas else None
            n = datas['labels_masks'][0] if 'labels_masks' in datas else None
        nl, nb = self.hparams['n_labels'], self.hparams['n_background']
        alpha, delta = self.hparams['ps_vae.alpha'], self.hparams['ps_vae.delta']
        beta, kl = self.beta_vals[self.curr_epoch], self.kl_anneal_vals[self.curr_epoch]
        x_hat, sample, mu, logvar, y_hat = self.forward(x, dataset=None, use_mean=False, eps=eps)
        t = {}
        t['loss_data_ll'] = gaussian_ll(x, x_hat, m)
        t['loss_label_ll'] = gaussian_ll(y, y_hat, n)
        t['loss_zs_kl'] = kl_div_to_std_normal(mu[:, :nl], logvar[:, :nl])
        t['loss_zu_mi'], t['loss_zu_tc'], t['loss_zu_dwkl'] = decomposed_kl(
            sample[:, nl + nb:], mu[:, nl + nb:], logvar[:, nl + nb:])
        total = (-t['loss_data_ll'] - alpha * t['loss_label_ll'] + t['loss_zs_kl'] + kl * t['loss_zu_mi']
                 + beta * t['loss_zu_tc'] + kl * t['loss_zu_dwkl'])
        if multi:
            t['loss_triplet'] = triplet_loss(self.TripletLoss, mu[:, nl:nl + nb], sessions)
            total = total + delta * t['loss_triplet']
        if accumulate_grad:
            total.backward()
        out = {'loss': total.item()}
        out.update({k: v.item() for k, v in t.items()})
        if not multi:
            out['loss_triplet'] = 0          # the reference keeps the zero it initialised the entry with
        n_dims = int(np.prod(x.shape[1:]))
        out['loss_data_mse'] = float(gaussian_ll_to_mse(out['loss_data_ll'], n_dims))
        y_np, yh_np = y.detach().cpu().numpy(), y_hat.detach().cpu().numpy()
        if n is not None:
            keep = n.detach().cpu().numpy() == 1
            r2 = r2_variance_weighted(y_np[keep], yh_np[keep])
        else:
            r2 = r2_variance_weighted(y_np, yh_np)
        out.update({'alpha': alpha, 'beta': beta, 'delta': delta, 'label_r2': r2})
        return out

    def get_predicted_labels(self, x, dataset=None, use_mean=True):
        z_s, _, _, logvar, _, _ = self.encoding(x, dataset=dataset)
        if not use_mean:
            z_s = reparameterize(z_s, logvar[:, :self.n_labels])
        return self.encoding.D(z_s)

    def _split(self, latents):
        nl, nb = self.hparams['n_labels'], self.hparams['n_background']
        return latents[:, :nl], latents[:, nl:nl + nb], latents[:, nl + nb:]

    def get_transformed_latents(self, inputs, dataset=None, as_numpy=True):
        """[labels-space supervised | background | unsupervised] from frames or latents (vaes.py:1105-1152)."""
        if not isinstance(inputs, torch.Tensor):
            inputs = torch.Tensor(inputs)
        inputs = inputs.to(self.encoding.D.weight.device)
        if inputs.dim() == 2:
            z_s, z_b, z_u = self._split(inputs)
        else:
            z_s, z_b, z_u = self.encoding(inputs, dataset=dataset)[:3]
        out = torch.cat([self.encoding.D(z_s), z_b, z_u], dim=1)
        return out.cpu().detach().numpy() if as_numpy else out

    def get_inverse_transformed_latents(self, inputs, dataset=None, as_numpy=True):
        """Inverse of the above for 2-D inputs (vaes.py:1154-1200)."""
        if not isinstance(inputs, torch.Tensor):
            inputs = torch.Tensor(inputs)
        if inputs.dim() != 2:
            raise NotImplementedError
        inputs = inputs.to(self.encoding.D.weight.device)
        y, z_b, z_u = self._split(inputs)
        z_s = torch.div(torch.sub(y, self.encoding.D.bias), self.encoding.D.weight)
        out = torch.cat([z_s, z_b, z_u], dim=1)
        return out.cpu().detach().numpy() if as_numpy else out

    def export_latents(self, data_gen, filename=None):
        """Latents of every trial as [z_s, z_b, z_u] (vaes.py:1202-1273).  Like the reference, the multi-session
        generator (whose training calls serve groups of trials from different sessions) is replaced by a
        single-session generator over the same sessions with every trial in the training split
        (``n_sessions_per_batch = 1``, ``train_frac = 1``, ``trial_splits = '1;0;0;0'``), so no trial is skipped."""
        from behavenet_b200.data.data_generator import PrefetchSessionsGenerator
        from behavenet_b200.fitting.eval import export_latents
        single = PrefetchSessionsGenerator(
            [ds.source for ds in data_gen.datasets], device=data_gen.device, rng_seed=0,
            trial_splits={'train_tr': 1, 'val_tr': 0, 'test_tr': 0, 'gap_tr': 0}, train_frac=1.0,
            raw_uint8=getattr(data_gen, 'raw_uint8', False), transforms=getattr(data_gen, 'transforms', None))
        try:
            return export_latents(single, self, filename=filename)
        finally:
            single.close()
