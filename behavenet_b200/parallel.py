"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The reference's only multi-GPU hook is an inert ``nn.DataParallel`` wrapper enabled by
``n_parallel_gpus`` (reference ``behavenet/models/base.py:106-116``, ``fitting/ae_grid_search.py:93-95``;
SURVEY.md section 2.1).  Here ``n_parallel_gpus = N`` means N ranks launched with torchrun: the
frame batch (CAE) or the trial axis (ARHMM) is partitioned contiguously over ranks and one
all-reduce(SUM) per step carries the flat gradient (CAE) or the sufficient statistics (ARHMM).
"""

import os

import torch

_STATE = {'group': None, 'enabled': False}


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return False
    if not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend=backend)
    _STATE['group'] = dist.group.WORLD
    _STATE['enabled'] = True
    return True


def enabled():
    return _STATE['enabled']


def group():
    return _STATE['group']


def world_size():
    if not _STATE['enabled']:
        return 1
    import torch.distributed as dist
    return dist.get_world_size(_STATE['group'])


def rank():
    if not _STATE['enabled']:
        return 0
    import torch.distributed as dist
    return dist.get_rank(_STATE['group'])


def shard_range(n, world=None, r=None):
    """Contiguous [begin, end) slice of ``n`` units owned by rank ``r`` (ceil split)."""
    world = world_size() if world is None else world
    r = rank() if r is None else r
    per = -(-n // world) if world > 0 else n
    return min(n, r * per), min(n, (r + 1) * per)


def all_reduce_sum(t):
    if _STATE['enabled']:
        import torch.distributed as dist
        dist.all_reduce(t, group=_STATE['group'])
    return t


def all_reduce_sum_async(t):
    """Start an all-reduce(SUM) and return its work handle (``.wait()`` orders the current stream after
    it).  With NCCL the collective runs on the communicator's own stream, after the work already enqueued
    on the current stream and concurrently with whatever is enqueued next."""
    import torch.distributed as dist
    return dist.all_reduce(t, group=_STATE['group'], async_op=True)


def overlap_enabled():
    """BN_DP_OVERLAP=0 issues the whole gradient as one all-reduce after the backward pass (A/B switch)."""
    return _STATE['enabled'] and os.environ.get('BN_DP_OVERLAP', '1') != '0'


def shutdown():
    import torch.distributed as dist
    if _STATE['enabled'] and dist.is_initialized():
        dist.destroy_process_group()
    _STATE['enabled'] = False
    _STATE['group'] = None
