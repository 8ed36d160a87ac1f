"""Determinism / batch-independence of one layer op: python scripts/diag_fprop_det.py side layer op
runs the op on 256 frames twice and on the first 32 frames alone; prints max |diff| (0 expected for a tile-local kernel)
and dumps the 256-frame result to /tmp/op_<tag>.pt for a cross-kernel comparison (BN_FPROP_HALO=0)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cae_oracle as co
from behavenet_b200 import _lib
from behavenet_b200.models import AE
side, layer, op = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
hp = co.make_hparams(1, 128, 128, 12)
model = AE(copy.deepcopy(hp)); model.load_state_dict(co.init_state_dict(hp, seed=0)); model.cuda()
drv, rt = model._driver, model._rt
params = model._kernel_params()
dev = torch.device('cuda', 0)
lib = _lib.lib()
g = torch.Generator().manual_seed(3)
a256 = torch.randn(256, 64, 64, 32, generator=g).cuda()
def run(a):
    n = a.shape[0]
    packed = drv.packed(rt, params, dev); ws = drv.workspace(rt, n, dev)
    out = torch.empty(n, 32, 32, 64, device=dev)
    _lib.check(lib.bn_cae_layer_op(drv.plan(dev), side, layer, op, n, a.data_ptr(), None, out.data_ptr(), drv.table(params),
                                   packed.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), 'op')
    torch.cuda.synchronize()
    return out
o1 = run(a256); o2 = run(a256); o3 = run(a256[:32].contiguous())
print('repeat diff %.3e   batch-32 vs batch-256 diff %.3e   max |out| %.3f' % (float((o1 - o2).abs().max()), float((o1[:32] - o3).abs().max()), float(o1.abs().max())))
tag = os.environ.get('BN_FPROP_HALO', '1')
torch.save(o1.cpu(), '/tmp/op_%s.pt' % tag)
if os.path.exists('/tmp/op_1.pt') and os.path.exists('/tmp/op_0.pt'):
    x, y = torch.load('/tmp/op_1.pt'), torch.load('/tmp/op_0.pt')
    d = (x - y).abs()
    print('halo vs im2col kernel: max diff %.3e  mean diff %.3e  (max |out| %.3f)' % (float(d.max()), float(d.mean()), float(y.abs().max())))
