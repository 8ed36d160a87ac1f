"""How does the tensor core read fp32 operands of a kind::tf32 MMA?  Single-layer experiment on B200.

Runs one layer op (bn_cae_layer_op, mode 1) on random fp32 data that is NOT TF32-exact and compares
with an fp64 CPU convolution whose activation operand is (a) untouched, (b) truncated to 10 mantissa
bits, (c) rounded to nearest (ties away), (d) rounded to nearest even; weights rounded rna as the
packer does.  The variant that lands at ~1e-6 is the hardware's behaviour.
"""

import copy
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from behavenet_b200 import _lib                      # noqa: E402
from behavenet_b200.models import AE                 # noqa: E402
from oracle import cae_oracle as co                  # noqa: E402


def rne(t):
    b = t.contiguous().view(torch.int32)
    lsb = (b >> 13) & 1
    return ((b + 0xFFF + lsb) & ~0x1FFF).view(torch.float32)


VARIANTS = {'none': lambda t: t, 'trunc': co.tf32_trunc, 'rna': co.tf32_rna, 'rne': rne}


def main():
    n = 16
    hp = co.make_hparams(1, 128, 128, 12)
    sd = co.init_state_dict(hp, seed=0)
    model = AE(copy.deepcopy(hp))
    model.load_state_dict(sd)
    model.cuda()
    lib = _lib.lib()
    lib.bn_set_tensor_core_mode(1)
    drv, rt = model._driver, model._rt
    params = model._kernel_params()
    dev = torch.device('cuda')
    packed = drv.packed(rt, params, dev)
    ws = drv.workspace(rt, n, dev)
    g = torch.Generator().manual_seed(0)

    def layer_op(side, layer, op, a, b_, out):
        _lib.check(lib.bn_cae_layer_op(drv.plan(dev), side, layer, op, n, a.data_ptr(), _lib.ptr(b_), out.data_ptr(),
                                       drv.table(params), packed.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), 'op')
        torch.cuda.synchronize()

    # encoder conv2 forward (igemm_tma): 64 -> 128 channels, 32x32 -> 16x16
    x = torch.randn(n, 64, 32, 32, generator=g)
    w = sd['encoding.encoder.conv2.weight']
    bias = sd['encoding.encoder.conv2.bias']
    out = torch.empty(n, 16, 16, 128, device=dev)
    layer_op(0, 2, 0, x.permute(0, 2, 3, 1).contiguous().cuda(), None, out)
    got = out.cpu().permute(0, 3, 1, 2).double()
    print('encoder conv2 forward (igemm_tma, TF32): max|err| / max|ref| against fp64 conv with the activation operand ...')
    for name, q in VARIANTS.items():
        ref = F.leaky_relu(F.conv2d(F.pad(q(x).double(), (1, 2, 1, 2)), co.tf32_rna(w).double(), bias.double(), stride=2), 0.05)
        print('  %-6s %.3e' % (name, float((got - ref).abs().max() / ref.abs().max())))
    # same with exact weights, to see the weight rounding
    ref = F.leaky_relu(F.conv2d(F.pad(co.tf32_trunc(x).double(), (1, 2, 1, 2)), w.double(), bias.double(), stride=2), 0.05)
    print('  trunc activations, UNROUNDED weights: %.3e' % float((got - ref).abs().max() / ref.abs().max()))

    # weight gradient of the same layer (wgrad_tma): both operands are activations
    dy = torch.randn(n, 128, 16, 16, generator=g)
    gout = torch.zeros_like(w).cuda()
    layer_op(0, 2, 2, x.permute(0, 2, 3, 1).contiguous().cuda(), dy.permute(0, 2, 3, 1).contiguous().cuda(), gout)
    got = gout.cpu().double()
    print('encoder conv2 weight gradient (wgrad_tma, TF32):')
    for name, q in VARIANTS.items():
        wv = w.double().clone().requires_grad_(True)
        y = F.conv2d(F.pad(q(x).double(), (1, 2, 1, 2)), wv, None, stride=2)
        ref, = torch.autograd.grad(y, wv, q(dy).double())
        print('  %-6s %.3e' % (name, float((got - ref).abs().max() / ref.abs().max())))

    # decoder convtranspose2 forward (halo kernel): 128 -> 64 channels, 16x16 -> 32x32
    z = torch.randn(n, 128, 16, 16, generator=g)
    wt = sd['decoding.decoder.convtranspose2.weight']
    bt = sd['decoding.decoder.convtranspose2.bias']
    out = torch.empty(n, 32, 32, 64, device=dev)
    layer_op(1, 2, 0, z.permute(0, 2, 3, 1).contiguous().cuda(), None, out)
    got = out.cpu().permute(0, 3, 1, 2).double()
    print('decoder convtranspose2 forward (dgrad_halo, TF32):')
    for name, q in VARIANTS.items():
        full = F.conv_transpose2d(q(z).double(), co.tf32_rna(wt).double(), bt.double(), stride=2)
        ref = F.leaky_relu(full[:, :, 1:-2, 1:-2], 0.05)
        print('  %-6s %.3e' % (name, float((got - ref).abs().max() / ref.abs().max())))

    # first layer forward (thin_fprop_tc)
    x0 = torch.rand(n, 1, 128, 128, generator=g)
    w0, b0 = sd['encoding.encoder.conv0.weight'], sd['encoding.encoder.conv0.bias']
    out = torch.empty(n, 64, 64, 32, device=dev)
    layer_op(0, 0, 0, x0.permute(0, 2, 3, 1).contiguous().cuda(), None, out)
    got = out.cpu().permute(0, 3, 1, 2).double()
    print('encoder conv0 forward (thin_fprop_tc, TF32):')
    for name, q in VARIANTS.items():
        ref = F.leaky_relu(F.conv2d(F.pad(q(x0).double(), (1, 2, 1, 2)), co.tf32_rna(w0).double(), b0.double(), stride=2), 0.05)
        print('  %-6s %.3e' % (name, float((got - ref).abs().max() / ref.abs().max())))


if __name__ == '__main__':
    main()
