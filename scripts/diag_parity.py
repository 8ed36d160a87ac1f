"""Measured deviation table of the TF32 (mode 1) path, per parameter, on a B200.

    python scripts/diag_parity.py [out.txt]

For each case: max|g - g_ref| / max|g_ref| of every parameter gradient of
  mode 0 vs the fp32 CPU oracle, mode 1 vs the fp32 CPU oracle (the TF32-vs-fp32 distance) and
  mode 1 vs the TF32-EMULATING CPU oracle (oracle/cae_oracle.py tf32_emulation: same operand rounding,
  so only accumulation order differs), with truncated and with rounded activations.
The bounds in tests/test_gpu_cae_fullsize.py are ~2x the numbers this prints (profiles/r02_parity.txt).
"""

import copy
import sys

import torch

sys.path.insert(0, '.')
from behavenet_b200 import _lib                      # noqa: E402
from behavenet_b200.models import AE, PSVAE          # noqa: E402
from oracle import cae_oracle as co                  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run_case(name, c, h, w, L, b, chunk, mc='ae', nl=0, out=sys.stdout):
    hp = co.make_hparams(c, h, w, L, mc, nl)
    sd = co.init_state_dict(hp, seed=1)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(b, c, h, w, generator=g)
    y = torch.randn(b, nl, generator=g) if nl else None
    eps = torch.randn(b, L, generator=g) if nl else None

    def oracle():
        if mc == 'ae':
            return co.ae_loss(sd, hp, x, None, chunk_size=chunk)
        return co.psvae_loss(sd, hp, x, y, eps, chunk_size=chunk)

    l32, g32 = oracle()
    with co.tf32_emulation('trunc'):
        lt, gt = oracle()
    with co.tf32_emulation('rna'):
        lr, gr = oracle()
    res = {}
    for mode in (0, 1):
        model = (AE if mc == 'ae' else PSVAE)(copy.deepcopy(hp))
        model.load_state_dict(sd)
        model.cuda()
        model.curr_epoch = 1
        _lib.lib().bn_set_tensor_core_mode(mode)
        data = {'images': x.cuda()[None]}
        kw = {}
        if nl:
            data['labels'] = y.cuda()[None]
            kw['eps'] = eps.cuda()
        lo = model.loss(data, chunk_size=chunk, **kw)
        res[mode] = (lo, {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None})
    _lib.lib().bn_set_tensor_core_mode(1)
    print('\n== %s  (%dx%dx%d, %d latents, batch %d, chunks of %d, %s)' % (name, h, w, c, L, b, chunk, mc), file=out)
    print('loss: oracle fp32 %.10g  emu-trunc %.10g  emu-rna %.10g  mode0 %.10g  mode1 %.10g'
          % (l32['loss'], lt['loss'], lr['loss'], res[0][0]['loss'], res[1][0]['loss']), file=out)
    print('%-44s %10s %10s %10s %10s' % ('param', 'm0-fp32', 'm1-fp32', 'm1-emuT', 'm1-emuR'), file=out)
    for k in g32:
        if k not in res[0][1]:
            continue
        print('%-44s %10.2e %10.2e %10.2e %10.2e' % (
            k, rel(res[0][1][k], g32[k]), rel(res[1][1][k], g32[k]), rel(res[1][1][k], gt[k]),
            rel(res[1][1][k], gr[k])), file=out)
    out.flush()


if __name__ == '__main__':
    out = open(sys.argv[1], 'w') if len(sys.argv) > 1 else sys.stdout
    run_case('C2 full', 1, 128, 128, 12, 256, 200, out=out)
    run_case('C2 small', 1, 128, 128, 12, 24, 16, out=out)
    run_case('C3 full', 2, 128, 128, 16, 512, 200, 'ps-vae', 4, out=out)
    run_case('integration geometry', 1, 64, 48, 6, 64, 50, out=out)
    run_case('C1', 1, 32, 32, 8, 32, 200, out=out)
