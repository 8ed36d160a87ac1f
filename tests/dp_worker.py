"""One rank of tests/test_gpu_cae_fullsize.py::test_data_parallel_ranks_reproduce_single_process.

Launched with RANK / WORLD_SIZE / MASTER_* in the environment; every rank uses cuda:0 and the gloo
backend (NCCL cannot place two ranks on one device), runs the data-parallel AE.loss and PSVAE.loss on the
same full-batch description and writes its loss dict and gradients; rank 0 also writes the
single-process result.
"""

import copy
import os
import sys

import torch

from behavenet_b200 import parallel
from behavenet_b200.models import AE, PSVAE
from oracle import cae_oracle as co


def run(cls, hp, sd, data, kw, dp):
    model = cls(copy.deepcopy(hp))
    model.load_state_dict(sd)
    model.cuda()
    model.curr_epoch = 1
    model.data_parallel = dp
    out = model.loss(data, **kw)
    grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
    return {'loss': {k: float(v) for k, v in out.items()}, 'grads': grads}


def main(outdir):
    rank = int(os.environ['RANK'])
    torch.cuda.set_device(0)
    assert parallel.init('gloo')
    from behavenet_b200 import _lib
    _lib.lib().bn_set_tensor_core_mode(0)
    g = torch.Generator().manual_seed(3)
    cases = {}
    hp = co.make_hparams(1, 64, 48, 6)
    x = torch.rand(100, 1, 64, 48, generator=g).cuda()
    m = (torch.rand(100, 1, 64, 48, generator=g) > 0.1).float().cuda()
    cases['ae'] = (AE, hp, co.init_state_dict(hp, seed=0), {'images': x[None], 'masks': m[None]}, {'chunk_size': 64})
    hp = co.make_hparams(2, 64, 64, 8, 'ps-vae', 3)
    x = torch.rand(300, 2, 64, 64, generator=g).cuda()
    y = torch.randn(300, 3, generator=g).cuda()
    eps = torch.randn(300, 8, generator=g).cuda()
    cases['psvae'] = (PSVAE, hp, co.init_state_dict(hp, seed=0), {'images': x[None], 'labels': y[None]},
                      {'chunk_size': 128, 'eps': eps})
    res = {k: run(*v, dp=True) for k, v in cases.items()}
    torch.save(res, os.path.join(outdir, 'rank%d.pt' % rank))
    if rank == 0:
        torch.save({k: run(*v, dp=False) for k, v in cases.items()}, os.path.join(outdir, 'single.pt'))
    torch.cuda.synchronize()
    parallel.shutdown()


if __name__ == '__main__':
    main(sys.argv[1])
