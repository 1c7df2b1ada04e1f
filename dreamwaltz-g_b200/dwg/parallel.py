"""View-level data parallelism of the SDS loop: one process per GPU, rank = view,
ONE all-reduce of the flattened parameter gradients per step (NCCL over NVLink on GPUs; the same
code runs on gloo for the CPU tests).  The reference has no distributed code at all."""
import numpy as np
import torch
import torch.distributed as dist


def rank_seed(base, rank):
    """Per-rank stream for camera / pose / timestep / noise draws (views are independent)."""
    return int(base) + int(rank)


def rank_rng(base, rank):
    return np.random.default_rng(rank_seed(base, rank))


def flatten_grads(params):
    ps = [p for p in params if p.grad is not None]
    return ps, torch.cat([p.grad.reshape(-1) for p in ps]) if ps else torch.zeros(0)


def allreduce_grads(params, average=False):
    """Sum (or average) the gradients of ``params`` across ranks with a single collective and
    scatter the result back into ``p.grad``.  Returns the number of bytes reduced."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    ps, flat = flatten_grads(params)
    if flat.numel() == 0:
        return 0
    dist.all_reduce(flat)
    if average:
        flat /= dist.get_world_size()
    o = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n
    return flat.numel() * flat.element_size()
