"""Kernel-level GPU tests: every tcgen05 (TF32) layer kernel against the fp32 CUDA-core kernel on
the SAME tensors, with data that is exactly representable in TF32 (multiples of 1/16 with small
magnitude) so that products are exact and the two paths may only differ by fp32 summation order
(tolerance 2e-5 relative to the output's max).  Shapes are the fat layers of config C2/C3."""

import copy
import ctypes as C

import pytest
import torch

from oracle import cae_oracle as co

pytestmark = pytest.mark.gpu


def _exact(shape, gen, scale=16, span=8):
    return (torch.randint(-span, span + 1, shape, generator=gen).float() / scale)


def _setup(n_ch=1, n=32, h=128, w=128):
    from behavenet_b200 import _lib
    from behavenet_b200.models import AE
    hp = co.make_hparams(n_ch, h, w, 12)
    model = AE(copy.deepcopy(hp))
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(_exact(p.shape, g, scale=64, span=8))
    model.cuda()
    drv, rt = model._driver, model._rt
    params = model._kernel_params()
    packed = drv.packed(rt, params, torch.device('cuda', 0))
    ws = drv.workspace(rt, n, torch.device('cuda', 0))
    return _lib, model, drv, params, packed, ws, hp


def _dims(hp, side, layer):
    if side == 0:
        cb = hp['ae_input_dim'][0] if layer == 0 else hp['ae_encoding_n_channels'][layer - 1]
        hb = hp['ae_input_dim'][1] if layer == 0 else hp['ae_encoding_y_dim'][layer - 1]
        wb = hp['ae_input_dim'][2] if layer == 0 else hp['ae_encoding_x_dim'][layer - 1]
        return (hb, wb, cb), (hp['ae_encoding_y_dim'][layer], hp['ae_encoding_x_dim'][layer],
                              hp['ae_encoding_n_channels'][layer])
    c0, h0, w0 = hp['ae_decoding_starting_dim']
    cs = c0 if layer == 0 else hp['ae_decoding_n_channels'][layer - 1]
    hs = h0 if layer == 0 else hp['ae_decoding_y_dim'][layer - 1]
    wsm = w0 if layer == 0 else hp['ae_decoding_x_dim'][layer - 1]
    return ((hp['ae_decoding_y_dim'][layer], hp['ae_decoding_x_dim'][layer],
             hp['ae_decoding_n_channels'][layer]), (hs, wsm, cs))


def _run(lib, drv, params, packed, ws, side, layer, op, n, a, b, out, mode):
    lib.lib().bn_set_tensor_core_mode(mode)
    lib.check(lib.lib().bn_cae_layer_op(
        drv.plan(a.device), side, layer, op, n, a.data_ptr(), None if b is None else b.data_ptr(),
        out.data_ptr(), drv.table(params), packed.data_ptr(), ws.data_ptr(), lib.stream_ptr()),
        'bn_cae_layer_op')
    torch.cuda.synchronize()


@pytest.mark.parametrize('geom', [(32, 128, 128), (7, 64, 48), (50, 96, 80), (256, 128, 128)])
@pytest.mark.parametrize('side,layer', [(0, 1), (0, 2), (0, 3), (0, 4), (1, 0), (1, 1), (1, 2), (1, 3)])
@pytest.mark.parametrize('op', [0, 1, 2])
def test_tensor_core_kernel_matches_cuda_core_kernel(side, layer, op, geom):
    """geom = (frames, H, W): the C2 shape, the integration-test shape with ragged tiles (rows and
    reduction lengths that are not multiples of the tile sizes), a mid-size odd shape, and the C2 shape at the
    BENCHMARKED batch of 256 frames (the split-K grids, persistent tile loops and per-CTA column-sum tables take
    their full-size paths there)."""
    n, h, w = geom
    lib, model, drv, params, packed, ws, hp = _setup(n=n, h=h, w=w)
    big, small = _dims(hp, side, layer)
    g = torch.Generator().manual_seed(100 * side + 10 * layer + op)
    xb = _exact((n,) + big, g).cuda()
    xs = _exact((n,) + small, g).cuda()
    fprop_form = (side == 0 and op == 0) or (side == 1 and op == 1)
    if op == 2:
        wshape = params[2 * layer].shape if side == 0 else params[2 * drv.n_layers + 6 + 2 * layer].shape
        outs = []
        for mode in (0, 1):
            out = torch.zeros(wshape, device='cuda')
            _run(lib, drv, params, packed, ws, side, layer, 2, n, xb, xs, out, mode)
            outs.append(out)
    else:
        src = xb if fprop_form else xs
        oshape = (n,) + (small if fprop_form else big)
        outs = []
        for mode in (0, 1):
            out = torch.full(oshape, float('nan'), device='cuda')
            _run(lib, drv, params, packed, ws, side, layer, op, n, src, None, out, mode)
            outs.append(out)
    ref, tc = outs
    assert torch.isfinite(tc).all()
    err = float((tc - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err


@pytest.mark.parametrize('n_ch', [1, 2])
@pytest.mark.parametrize('geom', [(32, 128, 128), (7, 64, 48)])
@pytest.mark.parametrize('side,layer,op', [(0, 0, 0), (1, 4, 1)])
def test_thin_layer_tensor_core_kernel_matches_cuda_core_kernel(side, layer, op, geom, n_ch):
    """First encoder layer forward / last decoder layer backward-data (1-2 image channels <-> 32 feature
    channels): the tcgen05 form of thin_fprop against the fp32 form on TF32-exact data."""
    n, h, w = geom
    lib, model, drv, params, packed, ws, hp = _setup(n_ch=n_ch, n=n, h=h, w=w)
    big, small = _dims(hp, side, layer)
    g = torch.Generator().manual_seed(7 + side)
    xb = _exact((n,) + big, g).cuda()
    outs = []
    for mode in (0, 1):
        out = torch.full((n,) + small, float('nan'), device='cuda')
        _run(lib, drv, params, packed, ws, side, layer, op, n, xb, None, out, mode)
        outs.append(out)
    ref, tc = outs
    assert torch.isfinite(tc).all()
    err = float((tc - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err


@pytest.mark.parametrize('n_ch', [1, 2, 4])
@pytest.mark.parametrize('geom', [(32, 128, 128), (7, 64, 48), (256, 128, 128)])
@pytest.mark.parametrize('side,layer', [(0, 0), (1, 4)])
def test_thin_layer_weight_gradient_tensor_core_matches_cuda_core(side, layer, geom, n_ch):
    """Weight gradient of the first encoder / last decoder layer: the tcgen05 kernel (pixels as the GEMM
    reduction, cae_thin_tc.cu) against the fp32 kernel on TF32-exact data, incl. ragged tiles (64x48 -> 32x24
    feature map: 24 columns in a 32-wide tile) and the benchmarked batch."""
    n, h, w = geom
    if n_ch == 4 and n == 256:
        pytest.skip('4-channel frames at the full batch add nothing over the 1/2-channel cases')
    lib, model, drv, params, packed, ws, hp = _setup(n_ch=n_ch, n=n, h=h, w=w)
    big, small = _dims(hp, side, layer)
    g = torch.Generator().manual_seed(31 + side)
    xb = _exact((n,) + big, g).cuda()
    xs = _exact((n,) + small, g).cuda()
    wshape = params[2 * layer].shape if side == 0 else params[2 * drv.n_layers + 6 + 2 * layer].shape
    outs = []
    for mode in (0, 1):
        out = torch.zeros(wshape, device='cuda')
        _run(lib, drv, params, packed, ws, side, layer, 2, n, xb, xs, out, mode)
        outs.append(out)
    ref, tc = outs
    assert torch.isfinite(tc).all()
    err = float((tc - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err


@pytest.mark.parametrize('n_ch', [1, 2, 4])
@pytest.mark.parametrize('geom', [(32, 128, 128), (7, 64, 48), (50, 96, 80), (256, 128, 128)])
def test_output_layer_forward_with_fused_loss_tensor_core_matches_cuda_core(geom, n_ch):
    """Last decoder layer forward + sigmoid + fused loss: the GEMM-over-all-taps + col2im kernel (cae_thin_tc.cu)
    against the fp32 kernel on TF32-exact data: x_hat, dL/d(pre-sigmoid) and the per-chunk sums of squared
    errors (two reference chunks at 256 frames)."""
    n, h, w = geom
    if n_ch == 4 and n == 256:
        pytest.skip('4-channel frames at the full batch add nothing over the 1/2-channel cases')
    lib, model, drv, params, packed, ws, hp = _setup(n_ch=n_ch, n=n, h=h, w=w)
    big, small = _dims(hp, 1, 4)
    g = torch.Generator().manual_seed(77)
    xs = _exact((n,) + small, g).cuda()
    target = torch.rand((n, n_ch, h, w), generator=g).cuda()
    tot = n * n_ch * h * w
    nchunks = (n + 199) // 200
    outs = []
    for mode in (0, 1):
        out = torch.full((2 * tot + 2 * nchunks,), float('nan'), device='cuda')
        _run(lib, drv, params, packed, ws, 1, 4, 3, n, xs, target, out, mode)
        outs.append(out)
    ref, tc = outs
    assert torch.isfinite(tc[:2 * tot]).all()
    assert float((tc[:tot] - ref[:tot]).abs().max()) < 2e-6                       # x_hat in (0, 1)
    scale = float(ref[tot:2 * tot].abs().max())
    assert float((tc[tot:2 * tot] - ref[tot:2 * tot]).abs().max()) < 2e-5 * scale
    sse_ref = ref[2 * tot:].view(torch.float64)
    sse_tc = tc[2 * tot:].view(torch.float64)
    assert float(((sse_tc - sse_ref).abs() / sse_ref).max()) < 1e-6


def test_one_cta_kernels_behind_the_switches_stay_green():
    """The CTA-pair kernels take the 32- / 64-channel layers by default; the one-CTA halo and im2col kernels they
    replaced remain the path for shapes the pair kernels do not cover (and the A/B baseline of profiles/).  Run the
    layer-kernel parity cases of two geometries with the pair kernels switched off (switches are read once per
    process)."""
    import os
    import subprocess
    import sys
    from tests.helpers import ROOT
    env = dict(os.environ, BN_FPROP_HALO='0', BN_HALO_PAIR='0', BN_HALO_PAIR64='0')
    r = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_kernels.py', '-x', '-q', '-m', 'gpu', '-k',
                        'test_tensor_core_kernel_matches_cuda_core_kernel and (geom1 or geom2)'],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and ' passed' in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
