"""Phase timestamps of the tcgen05 emission kernel (CTA 0): python scripts/emit_dbg.py   (BN_EMIT=1: one-stage kernel)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['BN_EMIT_DBG'] = '1'
import numpy as np, torch
from oracle import arhmm_oracle as ao
from behavenet_b200 import _lib
from behavenet_b200.ssm import HMM
p = ao.synth_params()
hmm = HMM(16, 12, observations='ar', observation_kwargs={'lags': 2})
hmm.init_state_distn.log_pi0, hmm.transitions.log_Ps = p.log_pi0, p.log_Ps
hmm.observations.As, hmm.observations.bs, hmm.observations.Sigmas = p.As, p.bs, p.Sigmas
X = ao.sample_batch(p, 2048, 1000, seed=0)
st = hmm._stage([X[i] for i in range(2048)])
for _ in range(4): hmm._run_estep(st, True)
torch.cuda.synchronize()
raw = C.CDLL(_lib.lib()._name); buf = (C.c_longlong * 256)(); raw.bn_debug_emit_times(buf)
t = np.array(list(buf)).reshape(32, 8)
if os.environ.get('BN_EMIT') == '1':
    names = ['top', 'psi_posted', 'prefetch', 'acc_ready', 'epi_done']
else:
    names = ['bld_top', 'bld_slot', 'bld_posted', 'mma_go', 'mma_issued', 'epi_acc', 'epi_done', 'bld_written']
t0 = t[0, 0]
for i in range(24):
    print(i, ' '.join('%s=%d' % (nm, t[i, j] - t0) for j, nm in enumerate(names)))
