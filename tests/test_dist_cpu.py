"""Multi-process host logic of the ray-sharded path on CPU (gloo, world_size 2): ray partition, flat-bucket gradient
all-reduce, scalar reductions -- the same code bench.py runs over NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from robir_b200 import dist as rdist
    r, w, _ = rdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = rdist.shard_rays(1000, rank, world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(8, 4)
    frozen = torch.nn.Linear(4, 4)
    for p in frozen.parameters():
        p.requires_grad_(False)
    x = torch.arange(1000 * 8, dtype=torch.float32).reshape(1000, 8) / 1000.0
    loss = lin(x[lo:hi]).pow(2).sum() / 1000.0        # global-N normalisation as in loss.py:41
    loss.backward()
    red = rdist.GradAllReducer(list(lin.parameters()) + list(frozen.parameters()))
    red()
    t = rdist.allreduce_min_scalar(torch.tensor([3.0 + rank]))
    mx = rdist.max_over_ranks(1.0 + rank, "cpu")
    sm = rdist.sum_over_ranks(1.0 + rank, "cpu")
    if rank == 0:
        torch.save(dict(grad=lin.weight.grad.clone(), span=(lo, hi), mn=float(t), mx=mx, sm=sm), out)
    dist.barrier()
    dist.destroy_process_group()


def test_ray_sharding_and_grad_allreduce(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    torch.manual_seed(0)
    lin = torch.nn.Linear(8, 4)
    x = torch.arange(1000 * 8, dtype=torch.float32).reshape(1000, 8) / 1000.0
    (lin(x).pow(2).sum() / 1000.0).backward()
    # all-reduce averages over ranks: sum of the two half-batch gradients / 2
    assert torch.allclose(res["grad"] * 2, lin.weight.grad, rtol=1e-5, atol=1e-6)
    assert res["span"] == (0, 500) and res["mn"] == 3.0 and res["mx"] == 2.0 and res["sm"] == 3.0


def test_shard_rays_covers_everything():
    from robir_b200.dist import shard_rays
    for n in (0, 1, 7, 1024, 640000):
        for w in (1, 2, 3, 8):
            spans = [shard_rays(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def _strong_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from robir_b200 import dist as rdist
    from robir_b200.loss import InvLoss
    from robir_b200.sg_render import kl_divergence
    rdist.init_from_env(backend="gloo")
    rdist.STRONG_SHARDING = True
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(5, 6) * 0.5)
    inp = torch.randn(11, 5)                         # 11 rows: ragged split 6 + 5
    lo, hi = rdist.shard_rays(11, rank, world)
    x = torch.sigmoid(inp[lo:hi] @ w)
    loss = kl_divergence(x, 0.01) + InvLoss.kl_divergence(0.05, inp[lo:hi] @ w)
    loss.backward()
    rdist.GradAllReducer([w], average=False)()
    if rank == 0:
        torch.save(dict(loss=loss.detach(), grad=w.grad.clone()), out)
    dist.barrier()
    dist.destroy_process_group()


def test_strong_sharding_batch_statistics(tmp_path):
    """The two batch means inside the losses (CESR supervise KL, utils/utils.py:14-17; latent KL, model/loss.py:75-79)
    over a batch that is split across two ranks: same value on every rank, parameter gradients add up to the
    full-batch gradient."""
    out = str(tmp_path / "r0.pt")
    mp.spawn(_strong_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    from robir_b200.loss import InvLoss
    from robir_b200.sg_render import kl_divergence
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(5, 6) * 0.5)
    inp = torch.randn(11, 5)
    loss = kl_divergence(torch.sigmoid(inp @ w), 0.01) + InvLoss.kl_divergence(0.05, inp @ w)
    loss.backward()
    assert abs(res["loss"].item() - loss.item()) < 1e-6
    assert torch.allclose(res["grad"], w.grad, rtol=1e-5, atol=1e-7)
