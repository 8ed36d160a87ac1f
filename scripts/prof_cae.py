"""Profiling driver: a few C2-shaped AE.loss steps (or ARHMM E-steps) for ncu."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cae_oracle as co
from behavenet_b200 import _lib

what = sys.argv[1] if len(sys.argv) > 1 else 'cae'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = int(os.environ.get('BN_TC', '1'))
_lib.lib().bn_set_tensor_core_mode(mode)
if what == 'cae':
    from behavenet_b200.models import AE
    hp = co.make_hparams(1, 128, 128, 12)
    model = AE(copy.deepcopy(hp)); model.load_state_dict(co.init_state_dict(hp, seed=0)); model.cuda()
    x = torch.rand(256, 1, 128, 128, generator=torch.Generator().manual_seed(0)).cuda()
    for _ in range(steps):
        model.zero_grad()
        out = model.loss({'images': x[None]})
    torch.cuda.synchronize()
    print(out)
elif what == 'psvae':
    from behavenet_b200.models import PSVAE
    hp = co.make_hparams(2, 128, 128, 16, 'ps-vae', 4)
    model = PSVAE(copy.deepcopy(hp)); model.load_state_dict(co.init_state_dict(hp, seed=0)); model.cuda()
    model.curr_epoch = 1
    g = torch.Generator().manual_seed(0)
    x = torch.rand(512, 2, 128, 128, generator=g).cuda()
    y = torch.randn(512, 4, generator=g).cuda()
    for _ in range(steps):
        model.zero_grad()
        out = model.loss({'images': x[None], 'labels': y[None]})
    torch.cuda.synchronize()
    print(out)
else:
    import numpy as np
    from oracle import arhmm_oracle as ao
    from behavenet_b200.ssm import HMM
    p = ao.synth_params()
    hmm = HMM(16, 12, observations='ar', observation_kwargs={'lags': 2})
    hmm.init_state_distn.log_pi0, hmm.transitions.log_Ps = p.log_pi0, p.log_Ps
    hmm.observations.As, hmm.observations.bs, hmm.observations.Sigmas = p.As, p.bs, p.Sigmas
    X = ao.sample_batch(p, 2048, 1000, seed=0)
    st = hmm._stage([X[i] for i in range(2048)])
    for _ in range(steps):
        Ez, Ezz, logZ = hmm._run_estep(st, True)
    torch.cuda.synchronize()
    print(float(logZ.sum()))
