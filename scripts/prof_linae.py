"""One warm + two profiled AE.loss calls of the linear autoencoder (2048 frames of 128 x 128, 12 latents) for an ncu launch list."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import cae_oracle as co
from behavenet_b200.models import AE
n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, int(sys.argv[2]) if len(sys.argv) > 2 else 12
hp = co.make_linear_hparams(1, 128, 128, L)
model = AE(copy.deepcopy(hp)); model.load_state_dict(co.init_linear_state_dict(hp, seed=0)); model.cuda()
x = torch.rand(n, 1, 128, 128, device='cuda')
for _ in range(3):
    model.loss({'images': x[None]})
torch.cuda.synchronize()
