"""View-level data parallelism of the SDS loop: one process per GPU, rank = view,
ONE all-reduce of the parameter gradients per step (NCCL over NVLink on GPUs; the same code runs on
gloo for the CPU tests).  The reference has no distributed code at all.

Gradients live in ONE persistent flat fp32 buffer (`GradBucket`): every ``p.grad`` is a view into it,
so the exchange is a single in-place collective with no flatten / copy-back traffic, the parameter
list is fixed (a parameter that received no gradient this step contributes zeros instead of changing
the buffer length -- ranks can never disagree on the layout), and the fused optimiser
(dwg/optim.py) walks the same buffer.
"""
import numpy as np
import torch
import torch.distributed as dist


def rank_seed(base, rank):
    """Per-rank stream for camera / pose / timestep / noise draws (views are independent)."""
    return int(base) + int(rank)


def rank_rng(base, rank):
    return np.random.default_rng(rank_seed(base, rank))


class GradBucket:
    """Flat gradient storage for a fixed parameter list.  ``zero()`` replaces ``p.grad = None``;
    autograd then accumulates in place into the views (capturable in a CUDA graph)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, 'no trainable parameters'
        dev, dt = self.params[0].device, self.params[0].dtype
        assert all(p.device == dev and p.dtype == dt for p in self.params), 'one device / dtype per bucket'
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4                      # 16-byte aligned segments (vectorised optimiser)
        self.flat = torch.zeros(n, device=dev, dtype=dt)
        self.attach()

    def attach(self):
        """(Re)install the views as ``p.grad`` (needed after anything set ``p.grad = None``)."""
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view_as(p)

    def zero(self):
        self.flat.zero_()
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + o * self.flat.element_size():
                p.grad = self.flat[o:o + p.numel()].view_as(p)

    def all_reduce(self, average=False):
        """In-place sum (or mean) over ranks; returns the number of bytes reduced (0 on a single rank)."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return 0
        dist.all_reduce(self.flat)
        if average:
            self.flat /= dist.get_world_size()
        return self.flat.numel() * self.flat.element_size()

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


def allreduce_grads(params, average=False):
    """Sum (or average) the gradients of ``params`` across ranks with a single collective.  The parameter list is
    FIXED: a parameter without a gradient contributes zeros (and receives the reduced value), so every rank reduces
    a buffer of the same length.  Prefer GradBucket (no flatten / copy-back); this is the stateless form.
    Returns the number of bytes reduced."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
    dist.all_reduce(flat)
    if average:
        flat /= dist.get_world_size()
    o = 0
    for p in ps:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[o:o + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n
    return flat.numel() * flat.element_size()
