"""Shard-sum vs unsharded gradient difference per parameter (tests/test_gpu_cae_fullsize.py), with / without BN_FPROP_HALO=0."""
import os, sys, copy
sys.path.insert(0, '/root/repo')
import torch
from tests.test_gpu_cae_fullsize import _model
from tests import test_gpu_cae_fullsize as T
from behavenet_b200 import parallel
from behavenet_b200.models import AE
import inspect
from oracle import cae_oracle as co
hp = co.make_hparams(1, 128, 128, 12)
c2 = dict(hp=hp, sd=co.init_state_dict(hp, seed=1), x=torch.rand(256, 1, 128, 128, generator=torch.Generator().manual_seed(5)))
x = c2['x'].cuda()
m = (torch.rand(x.shape, generator=torch.Generator().manual_seed(9)) > 0.1).float().cuda()
model = _model(AE, c2['hp'], c2['sd'], 1)
full = model.loss({'images': x[None], 'masks': m[None]})
gfull = {k: p.grad.clone() for k, p in model.named_parameters()}
for world in (2, 8):
    acc, scale = None, None
    for r in range(world):
        b, e = parallel.shard_range(x.shape[0], world, r)
        model.zero_grad()
        model.loss({'images': x[b:e][None], 'masks': m[b:e][None], 'shard': (b, x.shape[0])})
        gs = {k: p.grad.double() for k, p in model.named_parameters()}
        acc = gs if acc is None else {k: acc[k] + gs[k] for k in gs}
        mx = {k: float(v.abs().max()) for k, v in gs.items()}
        scale = mx if scale is None else {k: max(scale[k], mx[k]) for k in mx}
    errs = {k: float((acc[k] - gfull[k].double()).abs().max()) / max(scale[k], 1e-30) for k in acc}
    print(world, os.environ.get('BN_FPROP_HALO'), ' '.join('%s=%.4f' % (k.split('.')[-2][-5:] + k[-2:], v) for k, v in errs.items()))
