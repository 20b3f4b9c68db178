"""-m gpu parity of the BENCHMARKED configuration: BASELINE config 2 exactly as bench.py runs it --
M = 128 light SGs, S = 32, N = 1024 random pixels of the 800x800 view, stage-2 SDF radius 0.87, device-side randoms,
fixed-capacity batch, whole step as one CUDA graph -- against the CPU oracle ON THE SAME RANDOMS:

  (c) one replay of ``GraphedPBRStep`` (the thing bench.py times); its random tensors are read back after the replay;
  (b) the eager fixed-capacity step (static_shapes=True, forward + loss + backward) replaying those randoms;
  (a) the eager dynamic-shape step replaying those randoms (per-hit rows only, reference draw order);
  (o) the oracle (oracle/pipeline.py) on the same randoms: every forward key <= 1e-4 relative, the loss, and the
      gradients of all 19 trained tensors (train_pbr.py:104-105: light SGs, specular reflectance, the 16 tensors of the
      spec-BRDF auto-encoder, the tone-mapper's exposure).

(c) == (b) == (a) are compared tightly (same kernels, same numbers, different orchestration), (a) vs (o) with the
north-star tolerance.  Gradient tolerance: relative L2 <= 1e-3 and relative max <= 1e-2 per tensor for the tensor-core
engine -- see ``test_relu_flip_claim`` in test_gpu_parity.py for why a ReLU network cannot do better than that
tensor-by-tensor in fp32, and for the demonstration that with the oracle's own ReLU masks the same backward agrees to 1e-5.
"""
import numpy as np
import pytest
import torch

import pipeline as P
import robir_oracle as O
import tracers as T
from robir_b200 import synthetic

pytestmark = pytest.mark.gpu

N_RAYS, M_LOBES, SDF_RADIUS, SEED = 1024, 128, 0.87, 0          # bench.py's constants
REL = 1e-4


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(1.0, b.abs().max().item())


def grad_err(a, b):
    if a is None or b is None:
        assert (a is None or float(a.abs().max()) == 0.0) and (b is None or float(b.abs().max()) == 0.0)
        return 0.0, 0.0
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if float(b.abs().max()) == 0.0:
        assert float(a.abs().max()) == 0.0
        return 0.0, 0.0
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item(), \
        ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def trained_params(model):
    return list(model.gamma.parameters()) + list(model.envmap_material_network.parameters())


def trained_names(model):
    return ["gamma." + k for k, _ in model.gamma.named_parameters()] + \
        ["envmap_material_network." + k for k, _ in model.envmap_material_network.named_parameters()]


@pytest.fixture(scope="module")
def bench_setup():
    import robir_b200
    from robir_b200 import rng
    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=M_LOBES, sdf_radius=SDF_RADIUS)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=M_LOBES)))
    model.load_state_dict(sd, strict=True)
    model.cuda().train()
    model.generate()
    pix = synthetic.training_pixels(4242, n=N_RAYS)
    inp = synthetic.camera_inputs(pix)
    gen = torch.Generator().manual_seed(9)
    gt = torch.rand(1, N_RAYS, 3, generator=gen)
    yield sd, model, inp, gt
    rng.set_mode("cpu")
    model.static_shapes = False


def _load(model, sd):
    with torch.no_grad():
        cur = model.state_dict()
        for k, v in sd.items():
            cur[k].copy_(v)
    from robir_b200 import ops
    ops.invalidate_packed_weights()


def _static_tape_to_host_order(tape, n_hit):
    """Device-mode draws of the fixed-capacity step -> the 9 reference-order tensors over the hit rows
    (SURVEY.md A.4): [indir (N,64), brdf (N,32), normal (N,60), theta (M,S), phi (M,S), pairs theta (2N,8), pairs phi]."""
    assert len(tape) == 7, [tuple(t.shape) for t in tape]
    ind, brdf, nrm, th, ph, pt, pp = [t.detach().cpu().clone() for t in tape]
    N = ind.shape[0]
    return [ind[:n_hit], brdf[:n_hit], nrm[:n_hit], th, ph, pt[:n_hit], pp[:n_hit], pt[N:N + n_hit], pp[N:N + n_hit]]


def _static_tape_for_replay(tape):
    """The same numbers in the shapes the eager fixed-capacity step asks for in replay mode (it draws the BRDF-lobe
    pairs as four [N, 8] tensors)."""
    ind, brdf, nrm, th, ph, pt, pp = [t.detach().clone() for t in tape]
    N = ind.shape[0]
    return [ind, brdf, nrm, th, ph, pt[:N], pp[:N], pt[N:], pp[N:]]


def test_benchmarked_graph_step_matches_eager_modes_and_oracle(bench_setup):
    import robir_b200
    from robir_b200 import graph, ops, rng
    from robir_b200.loss import InvLoss, pbr_step_loss
    sd, model, inp, gt = bench_setup
    dev = torch.device("cuda")
    params, names = trained_params(model), trained_names(model)
    assert ops.ENGINE["vis"] == "tc" and ops.ENGINE["mlp"] == "tc"       # what bench.py runs
    torch.cuda.manual_seed(20260)                                        # the device-side draws of the captured step
    # ---------------------------------------------------------------- (c) the captured step, one replay
    rng.set_mode("device")
    loss_fn = InvLoss()
    opt = torch.optim.Adam(params, lr=5e-4, capturable=True, fused=True)
    step = graph.GraphedPBRStep(model, loss_fn, opt, N_RAYS, synthetic.camera_pose().to(dev),
                                synthetic.camera_intrinsics().to(dev), record_randoms=True)
    _load(model, sd)                                 # construction trained for a few steps: back to the bench weights
    loss_c = float(step(inp["uv"].to(dev), inp["object_mask"].to(dev), gt.to(dev)))
    torch.cuda.synchronize()
    grads_c = [None if p.grad is None else p.grad.detach().clone() for p in params]
    tape_c = [t.detach().clone() for t in step.random_tape]
    n_hit = int(step.hits)
    assert 0.3 * N_RAYS < n_hit < 0.8 * N_RAYS, n_hit
    # ---------------------------------------------------------------- (b) eager fixed-capacity step, same randoms
    _load(model, sd)
    dinp = {k: v.to(dev) for k, v in inp.items()}
    dinp["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(N_RAYS, 1)
    model.static_shapes, loss_fn.static_shapes = True, True
    for p in params:
        p.grad = None
    with rng.replay(_static_tape_for_replay(tape_c)):
        out_b = model(dinp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
    loss_b, _ = pbr_step_loss(model, loss_fn, out_b, {"rgb": gt.to(dev)})
    loss_b.backward()
    grads_b = [None if p.grad is None else p.grad.detach().clone() for p in params]
    assert int(out_b["network_object_mask"].sum()) == n_hit
    assert abs(float(loss_b) - loss_c) < 2e-6 * max(1.0, abs(loss_c)), (float(loss_b), loss_c)
    for k, gb, gc in zip(names, grads_b, grads_c):
        e2, ei = grad_err(gc, gb)
        assert e2 < 2e-5 and ei < 2e-4, ("graph vs eager-static", k, e2, ei)
    # ---------------------------------------------------------------- (a) eager dynamic-shape step, same randoms
    model.static_shapes, loss_fn.static_shapes = False, False
    host_tape = _static_tape_to_host_order(tape_c, n_hit)
    dinp["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(N_RAYS, 1)
    for p in params:
        p.grad = None
    with rng.replay(host_tape):
        out_a = model(dinp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
    loss_a, _ = pbr_step_loss(model, InvLoss(), out_a, {"rgb": gt.to(dev)})
    loss_a.backward()
    grads_a = [None if p.grad is None else p.grad.detach().clone() for p in params]
    assert torch.equal(out_a["network_object_mask"], out_b["network_object_mask"])
    for k in out_a:
        if out_a[k].dtype != torch.bool and k != "points":
            assert rel_err(out_b[k], out_a[k]) < 1e-5, ("static vs dynamic", k, rel_err(out_b[k], out_a[k]))
    assert abs(float(loss_a) - float(loss_b)) < 2e-6 * max(1.0, abs(float(loss_a)))
    for k, ga, gb in zip(names, grads_a, grads_b):
        e2, ei = grad_err(gb, ga)
        assert e2 < 2e-5 and ei < 2e-4, ("eager-static vs dynamic", k, e2, ei)
    # ---------------------------------------------------------------- (o) the oracle on the same tree and randoms
    torch.set_num_threads(max(1, min(32, torch.get_num_threads() * 4)))
    tree = T.OctreeOracle.__new__(T.OctreeOracle)
    for k, v in model.ray_tracer.sdf_octree.host_arrays().items():
        setattr(tree, k, v)
    tree.max_iter = -1
    sdo = {k: v.detach().clone() for k, v in sd.items()}
    for k in names:
        sdo[k].requires_grad_(True)
    oinp = dict(inp)
    oinp["hdr_shift"] = O.hdr_shift_as_input(sdo).expand(N_RAYS, 1)
    ref = P.idr_forward(sdo, oinp, lambda c, m, d: tree.trace(c, d), P.tape_to_rnd([("r", t) for t in host_tape]))
    assert torch.equal(ref["network_object_mask"], out_a["network_object_mask"].cpu()), "tracer mask differs"
    worst = {}
    for k, v in ref.items():
        if v.dtype == torch.bool or k not in out_a:
            continue
        assert out_a[k].shape == v.shape, k
        if k == "vis_shadow":
            # a lobe-weighted mean of per-sample visibilities behind a HARD predicate (the sample is culled when
            # cos(normal, direction) <= 0): one of the ~2.2 M samples sitting within an ulp of that boundary moves one
            # ray's value by 1/32 of a lobe weight (~2.4e-4).  Allow a handful of such rays, bound the rest as usual.
            d = (out_a[k].detach().float().cpu() - v).abs().amax(-1)
            off = d > REL
            assert int(off.sum()) <= 3 and float(d.max()) < 2e-3, ("forward vs oracle", k, int(off.sum()), float(d.max()))
            worst[k] = float(d[~off].max())
            continue
        worst[k] = rel_err(out_a[k], v)
        assert worst[k] < REL, ("forward vs oracle", k, worst[k])
    loss_o, _ = O.pbr_loss(sdo, ref, gt)
    assert abs(float(loss_o) - float(loss_a)) < 1e-5 * max(1.0, abs(float(loss_o))), (float(loss_o), float(loss_a))
    loss_o.backward()
    report = []
    for k, ga in zip(names, grads_a):
        go = sdo[k].grad
        if go is None:                     # gamma.{gamma, coef, ...}: not on the PBR graph in either implementation
            assert ga is None or float(ga.abs().max()) == 0.0, k
            continue
        e2, ei = grad_err(ga, go)
        report.append((k, e2, ei))
    assert len(report) >= 19, len(report)
    print("\nbench-config gradients vs oracle (rel L2, rel max):")
    for k, e2, ei in report:
        print("  %-70s %.2e %.2e" % (k, e2, ei))
    print("bench-config forward vs oracle (worst rel):", max(worst.values()), max(worst, key=worst.get))
    for k, e2, ei in report:
        assert e2 < 1e-3 and ei < 1e-2, ("gradient vs oracle", k, e2, ei)


def test_eager_forward_after_graph_replays_uses_current_weights(bench_setup):
    """the captured Adam update changes the trained weights without bumping tensor versions; an
    eager forward after N replays must see the CURRENT weights, i.e. equal a freshly built model loaded from the same
    state dict."""
    import robir_b200
    from robir_b200 import graph, ops, rng
    from robir_b200.loss import InvLoss
    sd, model, inp, gt = bench_setup
    dev = torch.device("cuda")
    _load(model, sd)
    rng.set_mode("device")
    params = trained_params(model)
    opt = torch.optim.Adam(params, lr=5e-3, capturable=True, fused=True)
    step = graph.GraphedPBRStep(model, InvLoss(), opt, N_RAYS, synthetic.camera_pose().to(dev),
                                synthetic.camera_intrinsics().to(dev))
    for _ in range(5):
        step(inp["uv"].to(dev), inp["object_mask"].to(dev), gt.to(dev))
    torch.cuda.synchronize()
    model.static_shapes = False
    fresh = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=M_LOBES)))
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, strict=True)
    fresh.cuda().train()
    fresh.ray_tracer.sdf_octree = model.ray_tracer.sdf_octree
    fresh.octree_ray_tracer.sdf_octree = model.ray_tracer.sdf_octree
    dinp = {k: v.to(dev) for k, v in inp.items()}
    dinp["hdr_shift"] = model.gamma.hdr_shift.as_input().detach().expand(N_RAYS, 1)
    rng.set_mode("cpu")
    outs = []
    for m in (model, fresh):
        torch.manual_seed(99)
        with torch.no_grad():
            outs.append(m(dinp, trainstage="Material", train_spec=True))
    for k in ("sg_rgb", "indir_rgb", "diffuse_albedo", "roughness"):
        assert rel_err(outs[0][k], outs[1][k]) < 1e-6, (k, rel_err(outs[0][k], outs[1][k]))
    # and the weights did move (otherwise the test proves nothing)
    moved = (model.envmap_material_network.lgtSGs.detach().cpu() - sd["envmap_material_network.lgtSGs"]).abs().max()
    assert float(moved) > 1e-4


@pytest.mark.parametrize("cur_iter", [300, 600, 1200])
def test_cesr_graph_step_matches_eager_dynamic(bench_setup, cur_iter):
    """bench.py --config c4: the CESR step (train_cesr.py:465-559,387-430) as ONE CUDA graph over the fixed-capacity batch
    (IDRNetwork._forward_static -> ClusteredAlbedoHook.get_sg_render_static; shadow_net / normal_net skip the row tiles
    beyond the device-side hit count, the two supervise means run over the valid rows) against the eager dynamic-shape
    step -- the one pinned to the reference's goldens in test_cesr_step_vs_golden -- ON THE SAME RANDOMS: loss and the
    gradients of all trained tensors (19 PBR + 2 x 27 of the weight-normed networks), in the warm-up (300), explore (600)
    and project (1200: the render loss trains normal_net through d render / d normal) phases."""
    from robir_b200 import cesr, graph, ops, rng
    from robir_b200.loss import InvLoss
    sd, model, inp, gt = bench_setup
    dev = torch.device("cuda")
    _load(model, sd)
    sh, nr = synthetic.cesr_state_dicts(SEED)
    shadow, normal = cesr.WnMLP(191, 2), cesr.WnMLP(63, 3)
    shadow.load_state_dict(sh)
    normal.load_state_dict(nr)
    hook = cesr.ClusteredAlbedoHook(model, shadow.to(dev), normal.to(dev), cur_iter=cur_iter)
    old_hook = model.__dict__.get("get_sg_render")
    model.get_sg_render = hook.get_sg_render
    params = trained_params(model) + hook.parameters()
    names = trained_names(model) + ["shadow_net." + k for k, _ in hook.shadow_net.named_parameters()] + \
        ["normal_net." + k for k, _ in hook.normal_net.named_parameters()]

    def reload():
        _load(model, sd)
        with torch.no_grad():
            for net, ref in ((hook.shadow_net, sh), (hook.normal_net, nr)):
                cur = net.state_dict()
                for k, v in ref.items():
                    cur[k].copy_(v)

    try:
        rng.set_mode("device")
        torch.cuda.manual_seed(20261 + cur_iter)
        loss_fn = InvLoss()
        opt = torch.optim.Adam(params, lr=5e-4, capturable=True, fused=True)
        step = graph.GraphedPBRStep(model, loss_fn, opt, N_RAYS, synthetic.camera_pose().to(dev),
                                    synthetic.camera_intrinsics().to(dev), record_randoms=True, hook=hook)
        reload()
        loss_c = float(step(inp["uv"].to(dev), inp["object_mask"].to(dev), gt.to(dev)))
        torch.cuda.synchronize()
        grads_c = [None if p.grad is None else p.grad.detach().clone() for p in params]
        tape = [t.detach().cpu().clone() for t in step.random_tape]
        n_hit = int(step.hits)
        assert 0.3 * N_RAYS < n_hit < 0.8 * N_RAYS, n_hit
        assert step.launches_per_step > 60
        # ------------------------------------------------------------ eager dynamic-shape step on the same randoms
        N = N_RAYS
        if cur_iter > 1000:                      # normal-gradient branch: two get_specular_visibility calls
            assert len(tape) == 9, [tuple(t.shape) for t in tape]
            host_tape = [tape[0][:n_hit], tape[1][:n_hit], tape[2][:n_hit], tape[3], tape[4]] + \
                [t[:n_hit] for t in tape[5:]]
        else:
            host_tape = _static_tape_to_host_order(tape, n_hit)
        reload()
        model.static_shapes, loss_fn.static_shapes = False, False
        dinp = {k: v.to(dev) for k, v in inp.items()}
        dinp["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(N, 1)
        for p in params:
            p.grad = None
        with rng.replay(host_tape):
            out_a = model(dinp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss_a, _ = hook.pbr_step(InvLoss(), out_a, {"rgb": gt.to(dev)})
        loss_a.backward()
        assert int(out_a["network_object_mask"].sum()) == n_hit
        grads_a = [None if p.grad is None else p.grad.detach().clone() for p in params]
        # ------------------------------------------------------------ eager fixed-capacity step, same randoms
        reload()
        model.static_shapes, loss_fn.static_shapes = True, True
        dinp["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(N, 1)
        if cur_iter > 1000:
            static_tape = [t.clone() for t in tape]
        else:
            static_tape = _static_tape_for_replay(tape)
        for p in params:
            p.grad = None
        with rng.replay(static_tape):
            out_b = model(dinp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss_b, _ = hook.pbr_step(loss_fn, out_b, {"rgb": gt.to(dev)})
        loss_b.backward()
        grads_b = [None if p.grad is None else p.grad.detach().clone() for p in params]
        print("\nCESR step @%d: loss graph %.8f eager-static %.8f eager-dynamic %.8f" %
              (cur_iter, loss_c, float(loss_b), float(loss_a)))
        rows, checked = [], 0
        for k, ga, gb, gc in zip(names, grads_a, grads_b, grads_c):
            if ga is None or float(ga.abs().max()) == 0.0:
                assert gc is None or float(gc.abs().max()) == 0.0, k
                assert gb is None or float(gb.abs().max()) == 0.0, k
                continue
            rows.append((k, grad_err(gc, gb), grad_err(gb, ga)))
            checked += 1
        for k, (c2, ci), (b2, bi) in rows:
            if max(c2, b2) > 2e-5:
                print("  %-75s graph/static %.1e %.1e   static/dynamic %.1e %.1e" % (k, c2, ci, b2, bi))
        assert abs(float(loss_b) - loss_c) < 2e-6 * max(1.0, abs(loss_c)), (float(loss_b), loss_c)
        assert abs(float(loss_a) - loss_c) < 1e-5 * max(1.0, abs(loss_c)), (float(loss_a), loss_c)
        for k, (c2, ci), (b2, bi) in rows:
            assert c2 < 2e-5 and ci < 2e-4, ("graph vs eager-static", cur_iter, k, c2, ci)
            assert b2 < 1e-4 and bi < 1e-3, ("eager-static vs dynamic", cur_iter, k, b2, bi)
        assert checked >= (54 if cur_iter <= 500 else 19 + 27), checked
    finally:
        rng.set_mode("cpu")
        model.static_shapes = False
        if old_hook is None:
            model.__dict__.pop("get_sg_render", None)
        else:
            model.get_sg_render = old_hook


def test_evaluation_mode_forward_vs_oracle(bench_setup):
    """bench.py --config c2e: the plot_to_disk path (training/train_pbr.py:235-311) -- model.eval(), is_training = False
    (visibility evaluated in testing mode, sg_render.py:148-160,231-243), forward only, tone-mapped prediction -- at the
    benchmarked model size (M = 128, S = 32) on a 1024-pixel chunk of raster-ordered pixels, against the oracle on the
    same randoms."""
    from robir_b200 import rng
    sd, model, inp, gt = bench_setup
    dev = torch.device("cuda")
    _load(model, sd)
    rng.set_mode("cpu")
    model.static_shapes = False
    N = 1024
    pix = torch.arange(N) + 800 * 380 + 100            # a raster chunk through the middle of the image (split_input order)
    einp = synthetic.camera_inputs(pix)
    dinp = {k: v.to(dev) for k, v in einp.items()}
    model.eval()
    model.is_training = False
    try:
        torch.manual_seed(77)
        with rng.record() as tape, torch.no_grad():
            dinp["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(N, 1)
            out = model(dinp, trainstage="Material", lin_diff=False, fun_spec=False, train_spec=True)
            ldr = model.gamma.hdr_shift.hdr2ldr(out["sg_rgb"] + out["indir_rgb"])
    finally:
        model.train()
        model.is_training = True
    n_hit = int(out["network_object_mask"].sum())
    assert 0.2 * N < n_hit < N, n_hit
    torch.set_num_threads(max(1, min(32, torch.get_num_threads() * 4)))
    tree = T.OctreeOracle.__new__(T.OctreeOracle)
    for k, v in model.ray_tracer.sdf_octree.host_arrays().items():
        setattr(tree, k, v)
    tree.max_iter = -1
    sdo = {k: v.detach().clone() for k, v in sd.items()}
    oinp = dict(einp)
    oinp["hdr_shift"] = O.hdr_shift_as_input(sdo).expand(N, 1)
    with torch.no_grad():
        ref = P.idr_forward(sdo, oinp, lambda c, m, d: tree.trace(c, d), P.tape_to_rnd([("r", t) for t in tape]),
                            is_training=False)
    assert torch.equal(ref["network_object_mask"], out["network_object_mask"].cpu())
    worst = {}
    for k, v in ref.items():
        if v.dtype == torch.bool or k not in out:
            continue
        if k == "vis_shadow":                      # hard culling predicate: see the training-mode test above
            d = (out[k].float().cpu() - v).abs().amax(-1)
            assert int((d > REL).sum()) <= 3 and float(d.max()) < 2e-3
            continue
        worst[k] = rel_err(out[k], v)
        assert worst[k] < REL, ("evaluation forward vs oracle", k, worst[k])
    ldr_ref = O.hdr2ldr(ref["sg_rgb"] + ref["indir_rgb"], O.hdr_shift_as_input(sdo))       # model/loss.py:61-64
    assert rel_err(ldr, ldr_ref) < REL
    print("\nevaluation-mode forward vs oracle (worst rel):", max(worst.values()), max(worst, key=worst.get))


def test_render_image_public_entry_point(bench_setup):
    """robir_b200.render_image (the plot_to_disk compute path as a library call): chunking does not change what is rendered
    -- the traced mask and the per-point material maps (no random draws) are identical for 1024- and 4096-pixel chunks, the
    shaded images agree up to the Monte-Carlo noise of the visibility samples -- and the model's mode flags are restored."""
    import robir_b200
    sd, model, inp, gt = bench_setup
    dev = torch.device("cuda")
    _load(model, sd)
    from robir_b200 import rng
    rng.set_mode("device")
    H = W = 96
    pose = synthetic.camera_pose().to(dev)
    K = synthetic.camera_intrinsics(H, W, 1111.1 * H / 800.0).to(dev)
    model.train()
    try:
        a = robir_b200.render_image(model, pose, K, H, W, chunk=1024)
        a = {k: v.clone() for k, v in a.items()}
        b = robir_b200.render_image(model, pose, K, H, W, chunk=4096)
    finally:
        rng.set_mode("cpu")
    assert model.training and model.is_training and not model.static_shapes
    assert set(a) == {"pred_rgb", "sg_rgb", "indir_rgb", "diffuse_albedo", "roughness", "vis_shadow", "network_object_mask"}
    m = a["network_object_mask"]
    assert torch.equal(m, b["network_object_mask"]) and 0.2 < float(m.float().mean()) < 0.95
    for k in ("diffuse_albedo", "roughness"):       # batch size selects the engine of a chain (FFMA / layer engine): 1e-4
        assert a[k].shape == (H * W, 3) and rel_err(a[k], b[k]) < REL, k
    for k in ("pred_rgb", "sg_rgb", "vis_shadow"):
        assert torch.isfinite(a[k]).all() and torch.isfinite(b[k]).all()
        assert float((a[k] - b[k]).abs().mean()) < 0.02, (k, float((a[k] - b[k]).abs().mean()))
    assert float(a["pred_rgb"][~m].min()) >= 0.0        # rays that miss: the reference's fill value through the tone-mapper


def test_pipelined_trace_graph_step_equals_plain_graph_step(bench_setup):
    """GraphedPBRStep(pipeline_trace=True): step i's graph walks batch i + 1 through the octree under its own loss /
    backward and step i + 1 starts from the finished trace.  Same seeds, same batches: losses, hit counts and the trained
    parameters after four steps equal those of the plain graphed step (the tracer depends on nothing that is trained)."""
    from robir_b200 import graph, rng
    from robir_b200.loss import InvLoss
    sd, model, inp, gt = bench_setup
    dev = torch.device("cuda")
    params = trained_params(model)
    batches = []
    for s in range(5):
        b = synthetic.camera_inputs(synthetic.training_pixels(900 + s, n=N_RAYS))
        batches.append((b["uv"].to(dev), b["object_mask"].to(dev),
                        torch.rand(1, N_RAYS, 3, generator=torch.Generator().manual_seed(s)).to(dev)))
    rng.set_mode("device")
    res = {}
    try:
        for pipe in (False, True):
            _load(model, sd)
            torch.cuda.manual_seed(4711)
            opt = torch.optim.Adam(params, lr=5e-4, capturable=True, fused=True)
            step = graph.GraphedPBRStep(model, InvLoss(), opt, N_RAYS, synthetic.camera_pose().to(dev),
                                        synthetic.camera_intrinsics().to(dev), pipeline_trace=pipe)
            _load(model, sd)
            for st in opt.state.values():                     # the construction trained for a few steps: reset Adam
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
            losses, hits = [], []
            for s in range(4):
                uv, om, g = batches[s]
                if pipe:
                    loss = step(uv, om, g, batches[s + 1][0], batches[s + 1][1])
                else:
                    loss = step(uv, om, g)
                losses.append(float(loss))
                hits.append(int(step.hits))
            torch.cuda.synchronize()
            res[pipe] = (losses, hits, [p.detach().clone() for p in params])
            del step
    finally:
        rng.set_mode("cpu")
        model.static_shapes = False
    (l0, h0, p0), (l1, h1, p1) = res[False], res[True]
    assert h0 == h1, (h0, h1)
    print("\nplain vs pipelined losses:", l0, l1)
    for a, b in zip(l0, l1):
        assert abs(a - b) < 2e-5 * max(1.0, abs(a)), (l0, l1)
    for a, b in zip(p0, p1):
        assert rel_err(b, a) < 1e-5
