"""Data-parallel plumbing: rays shard across ranks (weak scaling, as nerfplusplus/ddp_train_nerf.py
:363-365,430 — each rank draws its own N_rand rays), parameters are replicated, and ONE all-reduce
per step moves the flat gradient buffer [coarse MLP | fine MLP | camera] (4.8 MB fp32) over
NCCL/NVLink.  Unlike the reference (nerfplusplus/create_nerf.py:64-65) the camera gradients are
reduced too, so replicas stay identical.
"""
import torch
import torch.distributed as dist


class FlatGrads:
    """One contiguous fp32 buffer with a named view per parameter tensor."""

    def __init__(self, named_tensors, device):
        self.names = [n for n, _ in named_tensors]
        sizes = [t.numel() for _, t in named_tensors]
        self.flat = torch.zeros(sum(sizes), device=device, dtype=torch.float32)
        self.views, off = {}, 0
        for (n, t), s in zip(named_tensors, sizes):
            self.views[n] = self.flat[off:off + s].view(t.shape)
            off += s

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """Mean over ranks (what DDP does), one collective.  NCCL averages inside the all-reduce (ReduceOp.AVG): no
        separate division kernel on the critical path; gloo (CPU tests) sums and divides."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.flat.is_cuda and dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        return self.flat

    def assign_to(self, named_params):
        """Point each parameter's .grad at its view (no copies)."""
        for n, p in named_params:
            p.grad = self.views[n]


def shard_rays(n_total, rank, world_size):
    """Contiguous [lo, hi) pixel range of ``rank`` for inference sharding
    (nerfplusplus/ddp_train_nerf.py:144-183 splits H*W evenly; here the remainder is spread)."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
