import copy, sys, time
sys.path.insert(0, '/root/repo')
import torch
from oracle import cae_oracle as co
from behavenet_b200.models import AE
hp = co.make_hparams(1, 128, 128, 12)
model = AE(copy.deepcopy(hp)); model.load_state_dict(co.init_state_dict(hp, seed=0)); model.cuda()
for n in (2, 256):
    x = torch.rand(n, 1, 128, 128).cuda()
    for _ in range(5): model.loss({'images': x[None]})
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(50): model.loss({'images': x[None]})
    torch.cuda.synchronize()
    print('n=%d: %.1f us per loss() call (host wall clock)' % (n, (time.perf_counter() - t) / 50 * 1e6))
import cProfile, pstats
x = torch.rand(2, 1, 128, 128).cuda()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): model.loss({'images': x[None]})
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
