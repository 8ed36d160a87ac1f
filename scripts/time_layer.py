"""Time one layer op through bn_cae_layer_op: python scripts/time_layer.py side layer op [n]"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cae_oracle as co
from behavenet_b200 import _lib
from behavenet_b200.models import AE

side, layer, op = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 256
hp = co.make_hparams(1, 128, 128, 12)
model = AE(copy.deepcopy(hp)); model.load_state_dict(co.init_state_dict(hp, seed=0)); model.cuda()
drv, rt = model._driver, model._rt
params = model._kernel_params()
dev = torch.device('cuda', 0)
packed = drv.packed(rt, params, dev)
ws = drv.workspace(rt, n, dev)
if side == 0:
    cb = 1 if layer == 0 else hp['ae_encoding_n_channels'][layer - 1]
    hb = 128 if layer == 0 else hp['ae_encoding_y_dim'][layer - 1]; wb = 128 if layer == 0 else hp['ae_encoding_x_dim'][layer - 1]
    cs, hs, wsm = hp['ae_encoding_n_channels'][layer], hp['ae_encoding_y_dim'][layer], hp['ae_encoding_x_dim'][layer]
else:
    c0, h0, w0 = hp['ae_decoding_starting_dim']
    cs = c0 if layer == 0 else hp['ae_decoding_n_channels'][layer - 1]
    hs = h0 if layer == 0 else hp['ae_decoding_y_dim'][layer - 1]; wsm = w0 if layer == 0 else hp['ae_decoding_x_dim'][layer - 1]
    cb, hb, wb = hp['ae_decoding_n_channels'][layer], hp['ae_decoding_y_dim'][layer], hp['ae_decoding_x_dim'][layer]
big = torch.rand(n, hb, wb, cb, device=dev); small = torch.rand(n, hs, wsm, cs, device=dev)
fprop_form = (side == 0 and op == 0) or (side == 1 and op == 1)
lib = _lib.lib()
if op == 2:
    wi = 2 * layer if side == 0 else 2 * drv.n_layers + 6 + 2 * layer
    out = torch.zeros_like(params[wi]); a, b = big, small
else:
    a, b = (big, None) if fprop_form else (small, None)
    out = torch.empty_like(small if fprop_form else big)
def run():
    _lib.check(lib.bn_cae_layer_op(drv.plan(dev), side, layer, op, n, a.data_ptr(), None if b is None else b.data_ptr(),
                                   out.data_ptr(), drv.table(params), packed.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), 'op')
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); e1.synchronize()
macs = n * hs * wsm * cs * 25 * cb
us = e0.elapsed_time(e1) * 50
print('side %d layer %d op %d: %.1f us  %.1f TFLOP/s  BN_HALO_DBG=%s' % (side, layer, op, us, 2 * macs / us * 1e-6, os.environ.get('BN_HALO_DBG')))
