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


def _fake_outputs(model, A, lo, hi, N=24):
    """Synthetic model outputs of an N-ray batch whose rows [lo, hi) this rank holds; every differentiable entry depends
    on a parameter (A, the light SGs, the spec-BRDF encoder through the loss's own re-encode of the hit points)."""
    g = torch.Generator().manual_seed(3)
    feat = torch.rand(N, 4, generator=g)
    pts = torch.randn(N, 3, generator=g) * 0.3
    hit = torch.rand(N, generator=g) > 0.3
    om = torch.rand(N, generator=g) > 0.1
    nrm = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1)
    lgt = model.envmap_material_network.lgtSGs
    rows = slice(lo, hi)
    f = feat[rows]
    out = {"sg_rgb": torch.sigmoid(f @ A[:, :3]) * lgt[:, 4:].abs().mean(), "indir_rgb": torch.sigmoid(f @ A[:, 3:6]) * 0.1,
           "diffuse_albedo": torch.sigmoid(f @ A[:, 6:9]), "random_xi_diffuse_albedo": torch.sigmoid(f @ A[:, 9:12]),
           "roughness": torch.sigmoid(f @ A[:, 12:15]), "random_xi_roughness": torch.sigmoid(f @ A[:, 15:18]),
           "network_object_mask": hit[rows], "object_mask": om[rows], "surface_mask": hit[rows], "points": pts[rows],
           "normal_map": nrm[rows], "normals": nrm[rows] * 0.9}
    gt = {"rgb": torch.rand(1, N, 3, generator=g)[:, rows]}
    return out, gt


def _loss_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import robir_b200
    from robir_b200 import dist as rdist
    from robir_b200.loss import InvLoss, pbr_step_loss
    rdist.init_from_env(backend="gloo")
    rdist.STRONG_SHARDING = True
    torch.manual_seed(0)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=16)))
    A = torch.nn.Parameter(torch.randn(4, 18) * 0.5)
    lo, hi = rdist.shard_rays(24, rank, world)
    mo, gt = _fake_outputs(model, A, lo, hi)
    loss, _ = pbr_step_loss(model, InvLoss(), mo, gt)
    loss.backward()
    params = [A, model.envmap_material_network.lgtSGs, model.gamma.hdr_shift.adapt_illum] + \
        list(model.envmap_material_network.spec_brdf_encoder_layer.brdf_encoder_layer.parameters())
    rdist.GradAllReducer(params, average=False)()
    if rank == 0:
        torch.save(dict(grads=[p.grad.clone() for p in params]), out)
    dist.barrier()
    dist.destroy_process_group()


def test_strong_sharding_full_pbr_step_loss_gradient(tmp_path):
    """the summed per-rank gradients of the WHOLE pbr_step_loss (image term / N, latent-smooth means, KL batch
    mean, white-light regulariser) over a batch split across two ranks equal the single-rank full-batch gradient."""
    out = str(tmp_path / "r0.pt")
    mp.spawn(_loss_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    import robir_b200
    from robir_b200.loss import InvLoss, pbr_step_loss
    torch.manual_seed(0)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=16)))
    A = torch.nn.Parameter(torch.randn(4, 18) * 0.5)
    mo, gt = _fake_outputs(model, A, 0, 24)
    loss, _ = pbr_step_loss(model, InvLoss(), mo, gt)
    loss.backward()
    params = [A, model.envmap_material_network.lgtSGs, model.gamma.hdr_shift.adapt_illum] + \
        list(model.envmap_material_network.spec_brdf_encoder_layer.brdf_encoder_layer.parameters())
    assert len(res["grads"]) == len(params)
    for g, p in zip(res["grads"], params):
        assert p.grad is not None and p.grad.abs().max() > 0
        assert torch.allclose(g, p.grad, rtol=2e-4, atol=1e-7), (g - p.grad).abs().max()
