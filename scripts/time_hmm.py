"""Time the C4 E-step pieces with CUDA events: python scripts/time_hmm.py [trials]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import arhmm_oracle as ao
from behavenet_b200 import _lib
from behavenet_b200.ssm import HMM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
p = ao.synth_params()
hmm = HMM(16, 12, observations='ar', observation_kwargs={'lags': 2})
hmm.init_state_distn.log_pi0, hmm.transitions.log_Ps = p.log_pi0, p.log_Ps
hmm.observations.As, hmm.observations.bs, hmm.observations.Sigmas = p.As, p.bs, p.Sigmas
X = ao.sample_batch(p, n, 1000, seed=0)
st = hmm._stage([X[i] for i in range(n)])

def timeit(fn, it=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

res = {}
for mode in (0, 1):
    _lib.lib().bn_set_tensor_core_mode(mode)
    res['estep_mode%d_us' % mode] = timeit(lambda: hmm._run_estep(st, True))
    res['loglik_mode%d_us' % mode] = timeit(lambda: hmm._run_estep(st, False))
    Ez, Ezz, logZ = hmm._run_estep(st, True)
    res['logZ_mode%d' % mode] = float(logZ.sum())
print(res)
