"""Diagnostic (GPU box): per-parameter gradient error of our two compute modes and of eager
PyTorch-on-CUDA (cuDNN, TF32 allowed = torch default) against the fp64 CPU oracle."""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cae_oracle as co
from behavenet_b200 import _lib
from behavenet_b200.models import AE

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
W = int(sys.argv[3]) if len(sys.argv) > 3 else 128
CH = int(sys.argv[4]) if len(sys.argv) > 4 else 16
hp = co.make_hparams(1, H, W, 12)
sd = co.init_state_dict(hp, seed=1)
x = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(5))
sd64 = {k: v.double() for k, v in sd.items()}
l64, g64 = co.ae_loss(sd64, hp, x.double(), None, chunk_size=CH)
l32, g32 = co.ae_loss(sd, hp, x, None, chunk_size=CH)

def rel(a, b):
    return float((a.double().cpu() - b).abs().max() / b.abs().max())

res = {}
res['cpu_fp32'] = {k: rel(g32[k], g64[k]) for k in g64}
for mode in (0, 1):
    model = AE(copy.deepcopy(hp)); model.load_state_dict(sd); model.cuda()
    _lib.lib().bn_set_tensor_core_mode(mode)
    out = model.loss({'images': x.cuda()[None]}, chunk_size=CH)
    res['ours_mode%d' % mode] = {k: rel(p.grad, g64[k]) for k, p in model.named_parameters()}
    print('mode', mode, 'loss', out['loss'], 'ref', l64['loss'])
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    sdc = {k: v.cuda() for k, v in sd.items()}
    l, g = co.ae_loss(sdc, hp, x.cuda(), None, chunk_size=CH)
    res['torch_cuda_tf32=%s' % tf32] = {k: rel(g[k], g64[k]) for k in g64}
keys = list(g64)
print('%-45s' % 'param' + ''.join('%22s' % n for n in res))
for k in keys:
    print('%-45s' % k + ''.join('%22.2e' % res[n][k] for n in res))
