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
    """Contiguous slice [lo, hi) of the ray axis owned by ``rank`` (uv order).  For exact strong
    sharding pass the whole batch to ``IDRNetwork.forward`` with ``input["shard"] = shard_rays(...)``: every rank then walks
    the full batch through the octree (its lock-step sample count is a batch-level quantity) and shades its slice."""
    per = (n_rays + world - 1) // world
    lo = min(rank * per, n_rays)
    return lo, min(lo + per, n_rays)


class GradAllReducer:
    """Flat-bucket all-reduce of the gradients of ``params``: one gather kernel, ONE NCCL call on the current stream
    (ReduceOp.AVG does the division inside the collective), one multi-tensor scatter.  Everything is stream-ordered
    device work with static shapes (``GraphedPBRStep`` can capture it inside the step's CUDA graph).  average=True (weak scaling:
    every rank steps on its own batch; the mean of the per-batch gradients is what a K-times larger batch would give);
    average=False is the plain sum that strong sharding of ONE batch needs: under ``STRONG_SHARDING`` the loss code
    normalises every per-ray term by the GLOBAL ray / hit counts (``global_count`` / ``global_mean``), shares the
    parameter-only terms out over the ranks (``param_only``) and takes the batch means of both KL terms over all ranks
    (``batch_mean_rows``), so the summed per-rank gradients equal the full-batch gradient (2-rank test:
    tests/test_dist_cpu.py)."""

    def __init__(self, params, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.average = average

    def __call__(self):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return
        grads = [p.grad for p in ps]
        flat = torch.cat([g.reshape(-1) for g in grads])              # one gather kernel
        if self.average and dist.get_backend() == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            if self.average:
                flat.div_(dist.get_world_size())
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(grads, views)                            # one scatter kernel


# ----------------------------------------------------------------------------------------------------------------------
# batch statistics for bit-faithful STRONG sharding of one batch: besides the sampler's minimum
# (allreduce_min_scalar) the losses contain two means over the batch axis -- the KL sparsity term of the spec-BRDF latent
# (model/loss.py:75-79) and the CESR supervise term (utils/utils.py:14-17 through sg_render.py:397-403).
# ----------------------------------------------------------------------------------------------------------------------
STRONG_SHARDING = False     # set True when the ranks render slices of the SAME batch (bench.py is weak scaling: False)


class _AllReduceSum(torch.autograd.Function):
    """S = sum over ranks of t.  Every rank evaluates the same loss L(S), so dL/dt_local = L'(S) = the incoming
    gradient; the parameter gradients of the ranks then add up to the full-batch gradient (GradAllReducer(average=False))."""

    @staticmethod
    def forward(ctx, t):
        out = t.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out

    @staticmethod
    def backward(ctx, g):
        return g


def batch_mean_rows(x):
    """torch.mean(x, 0) over the rows of the whole batch: local mean unless STRONG_SHARDING is on and a process group with
    more than one rank exists, in which case row sums and row counts are all-reduced (differentiable)."""
    if not (STRONG_SHARDING and dist.is_initialized() and dist.get_world_size() > 1):
        return torch.mean(x, 0)
    n = torch.tensor([float(x.shape[0])], dtype=x.dtype, device=x.device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return _AllReduceSum.apply(x.sum(0)) / n


def _strong():
    return STRONG_SHARDING and dist.is_initialized() and dist.get_world_size() > 1


def global_count(n, like):
    """Number of rows of the whole batch given the local count ``n`` (python number): model/loss.py:41 divides the
    image loss by ``object_mask.shape[0]`` of the full batch."""
    if not _strong():
        return float(n)
    t = torch.tensor([float(n)], dtype=torch.float64, device=like.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def global_mean(x):
    """x.mean() over the elements of the whole batch (all ranks hold slices with the same trailing shape)."""
    if not _strong():
        return x.mean()
    n = torch.tensor([float(x.numel())], dtype=x.dtype, device=x.device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return _AllReduceSum.apply(x.sum().reshape(1))[0] / n[0]


def param_only(term):
    """A loss term that depends on parameters only (white-light regulariser, ...): every rank evaluates it, the summed
    gradient must count it once."""
    return term / dist.get_world_size() if _strong() else term


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
