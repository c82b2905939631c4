"""Multi-GPU plumbing: one process per GPU (torchrun).  The hot path has no data-path collective -- catalogues are
independent, every rank processes its own -- so torch.distributed is only used for the barrier and for reducing the
timing / throughput bookkeeping (max over ranks, total units).  Works with NCCL (GPU) and gloo (CPU tests)."""
import os

import torch
import torch.distributed as dist


def rank_info():
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0)))


def init(backend=None, device=None):
    rank, world, _ = rank_info()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {'device_id': device} if (backend == 'nccl' and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(values, device='cpu'):
    """Element-wise max of a list of floats over all ranks (the bench's max-over-ranks timing)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device='cpu'):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def catalogue_seed(base, rank):
    """Every rank works on its own catalogue: distinct, reproducible seeds."""
    return int(base) + int(rank)


def seconds_per_catalogue(elapsed_s_max, steps, world):
    """Whole-job metric of bench.py: all ranks together processed steps*world catalogues in elapsed_s_max."""
    return elapsed_s_max / float(steps * world)
