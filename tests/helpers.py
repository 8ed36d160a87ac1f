"""Shared helpers for the parity tests."""

import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def golden_compare(gold, key, actual, rtol, atol, check_sum=True):
    """Compare ``actual`` with a golden entry stored whole or as (strided sample, checksums)."""
    a = actual.detach().cpu().numpy() if isinstance(actual, torch.Tensor) else np.asarray(actual)
    if key in gold:
        np.testing.assert_allclose(a, gold[key], rtol=rtol, atol=atol, err_msg=key)
        return
    idx, val = gold[key + '#idx'], gold[key + '#val']
    assert tuple(gold[key + '#shape']) == a.shape, key
    flat = a.reshape(-1)
    np.testing.assert_allclose(flat[idx], val, rtol=rtol, atol=atol, err_msg=key)
    if not check_sum:
        return
    s, sabs = gold[key + '#sum']
    n = flat.size
    assert abs(flat.astype(np.float64).sum() - s) <= atol * n + rtol * sabs, key
    assert abs(np.abs(flat.astype(np.float64)).sum() - sabs) <= atol * n + rtol * sabs, key


def synth_inputs(c, h, w, L, b, n_labels=0, seed=1234, variational=False):
    """Same draws as oracle/gen_golden.py::synth_inputs."""
    g = torch.Generator().manual_seed(seed)
    out = {'x': torch.rand(b, c, h, w, generator=g)}
    if n_labels:
        out['labels'] = torch.randn(b, n_labels, generator=g)
        out['eps'] = torch.randn(b, L, generator=g)
    elif variational:
        out['eps'] = torch.randn(b, L, generator=g)
    out['masks'] = (torch.rand(b, c, h, w, generator=g) > 0.1).float()
    return out


def synth_cond_inputs(c, h, w, b, n_labels, seed=1234, n_latents=0):
    """Inputs of the label-conditioned models (cond-ae, cond-ae-msp); oracle/gen_golden.py uses this too.
    labels_2d: one one-hot image per (x, y) label pair, as data/transforms.py MakeOneHot2D produces."""
    g = torch.Generator().manual_seed(seed)
    out = {'x': torch.rand(b, c, h, w, generator=g), 'labels': torch.randn(b, n_labels, generator=g)}
    n2 = n_labels // 2
    ys = torch.randint(0, h, (b, n2), generator=g)
    xs = torch.randint(0, w, (b, n2), generator=g)
    l2d = torch.zeros(b, n2, h, w)
    l2d[torch.arange(b)[:, None], torch.arange(n2)[None, :], ys, xs] = 1.0
    out['labels_2d'] = l2d
    out['masks'] = (torch.rand(b, c, h, w, generator=g) > 0.1).float()
    if n_latents:
        out['eps'] = torch.randn(b, n_latents, generator=g)
    return out


def rel_err(a, b):
    """max |a - b| / max |b|."""
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
