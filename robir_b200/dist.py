"""Multi-GPU plumbing: rays are independent units, so every rank renders its own slice / batch
with a full replica of the weights and the octree; the only data-path collective is one NCCL all-reduce of the flattened
gradient of the trained parameters after backward (+ optional scalar reductions for bit-faithful strong sharding)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """One process per GPU, launched by torchrun; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_rays(n_rays, rank, world):
    """Contiguous slice [lo, hi) of the ray axis owned by ``rank`` (uv order)."""
    per = (n_rays + world - 1) // world
    lo = min(rank * per, n_rays)
    return lo, min(lo + per, n_rays)


class GradAllReducer:
    """Flat-bucket all-reduce(SUM)/world of the gradients of ``params`` (one NCCL call on the current stream)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)

    def __call__(self):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return
        grads = [p.grad for p in ps]
        flat = torch.cat([g.reshape(-1) for g in grads])              # one gather kernel
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(grads, views)                            # one scatter kernel


def allreduce_min_scalar(t):
    """Batch-coupled scalar of the specular sampler (model/sg_render.py:220-222) for bit-faithful ray sharding."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t


def max_over_ranks(x, device):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
