"""Role timestamps of the fprop halo pair kernel (CTA 0, first 8 tiles): python scripts/fprop_dbg.py"""
import copy, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('BN_HALO_DBG', '1')
import numpy as np, torch
from oracle import cae_oracle as co
from behavenet_b200 import _lib
from behavenet_b200.models import AE
n = 256
hp = co.make_hparams(1, 128, 128, 12)
model = AE(copy.deepcopy(hp)); model.load_state_dict(co.init_state_dict(hp, seed=0)); model.cuda()
drv, rt = model._driver, model._rt
params = model._kernel_params()
dev = torch.device('cuda', 0)
packed = drv.packed(rt, params, dev); ws = drv.workspace(rt, n, dev)
a = torch.rand(n, 64, 64, 32, device=dev); out = torch.empty(n, 32, 32, 64, device=dev)
lib = _lib.lib()
def run():
    _lib.check(lib.bn_cae_layer_op(drv.plan(dev), 0, 1, 0, n, a.data_ptr(), None, out.data_ptr(), drv.table(params),
                                   packed.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), 'op')
for _ in range(5): run()
torch.cuda.synchronize()
raw = C.CDLL(lib._name); buf = (C.c_longlong * 64)(); raw.bn_debug_halo_times(buf)
t = np.array(list(buf)).reshape(8, 8); t0 = t[0, 0]
names = ['mma_go', 'pl0', 'pl1', 'pl2', 'pl3', 'mma_issued', 'epi_acc', 'epi_done']
for i in range(8):
    print(i, ' '.join('%s=%d' % (nm, t[i, j] - t0) for j, nm in enumerate(names)))
