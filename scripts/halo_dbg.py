"""Per-role phase timestamps of the persistent halo kernel (CTA 0): BN_HALO_DBG=1 python scripts/halo_dbg.py"""
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
a = torch.rand(n, 32, 32, 64, device=dev); out = torch.empty(n, 64, 64, 32, device=dev)
lib = _lib.lib()
def run():
    _lib.check(lib.bn_cae_layer_op(drv.plan(dev), 1, 3, 0, n, a.data_ptr(), None, out.data_ptr(), drv.table(params),
                                   packed.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), 'op')
for _ in range(5): run()
torch.cuda.synchronize()
raw = C.CDLL(lib._name); buf = (C.c_longlong * 64)(); raw.bn_debug_halo_times(buf)
t = np.array(list(buf)).reshape(8, 8); t0 = t[0, 0]
names = ['mma_top', 'acc_free', 'mma_issued', 'epi_wait', 'acc_ready', 'epi_done']
print('dbg', os.environ['BN_HALO_DBG'], 'per tile (tiles 2..6): mma phase us %.1f  epilogue us %.1f  tile period us %.1f' % (np.mean(t[2:7, 2] - t[2:7, 1]) / 1e3, np.mean(t[2:7, 5] - t[2:7, 4]) / 1e3, np.mean(np.diff(t[1:7, 5])) / 1e3))
