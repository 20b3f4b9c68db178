"""-m gpu parity tests: the CUDA product (through the C ABI) against the CPU oracle on identical seeded inputs, and
against the golden vectors generated from the unmodified reference.  Tolerances follow BASELINE.json north_star:
<= 1e-4 relative for fp32 outputs (rel = |a-b| / max(|b|_max, 1)); tracer masks must agree exactly on the same tree."""
import os

import numpy as np
import pytest
import torch

import pipeline as P
import robir_oracle as O
import tracers as T
from robir_b200 import synthetic

pytestmark = pytest.mark.gpu

REL = 1e-4


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(1.0, b.abs().max().item())


_ERR_LOG = os.environ.get("ROBIR_ERR_LOG")


def grad_close(a, b, l2=3e-3, linf=3e-2, engine="ffma"):
    """Gradients that pass through the ReLU visibility MLP are only piecewise continuous: a hidden unit whose
    pre-activation is ~0 can take a different sign on the GPU than in the CPU oracle, which changes a few entries by
    O(weight x upstream) while everything else agrees to ~1e-6.  How often that happens is set by the relative error of
    the pre-activations: ~1e-7 for fp32 and for the tensor-core engine's scaled fp16 hi/lo split (22 mantissa bits per
    operand; no flip in 2e7 units in tools/vis_numerics_study.py), ~1e-5 for the bf16 hi/lo split of round 1 (26 flips,
    gradient rel L2 2e-3; with the oracle's masks forced the same backward agrees to 7e-6 -- test_numerics_study.py).
    Both engines are therefore held to the same bounds now."""
    if engine == "tc_bf16":      # the per-layer engine of the 512-wide chains still splits into bf16 hi/lo (16 bits)
        l2, linf = 4 * l2, 4 * linf
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = max(b.abs().max().item(), 1e-12)
    e2 = (a - b).norm().item() / max(b.norm().item(), 1e-12)
    ei = (a - b).abs().max().item() / scale
    if _ERR_LOG:
        import inspect
        with open(_ERR_LOG, "a") as f:
            f.write("%-50s %-5s relL2 %.3e relmax %.3e (bounds %.1e %.1e)\n" %
                    (inspect.stack()[1].function, engine, e2, ei, l2, linf))
    assert torch.isfinite(a).all(), "non-finite gradient"
    assert e2 < l2 and ei < linf, "gradient mismatch: rel L2 %.3e (< %.1e), rel max %.3e (< %.1e)" % (e2, l2, ei, linf)


def ops_engine_mlp():
    from robir_b200 import ops
    return "tc_bf16" if ops.ENGINE["mlp"] == "tc" else ops.ENGINE["mlp"]


@pytest.fixture(scope="module")
def model16(synth_sd16):
    import robir_b200
    m = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=16)))
    m.load_state_dict(synth_sd16, strict=True)
    m.cuda().train()
    return m


def test_extension_is_loaded():
    from robir_b200 import _lib
    assert _lib.lib().robir_abi_version() == 1
    assert _lib.sm_count() >= 100


@pytest.fixture(params=["ffma", "tc"])
def engine(request):
    """Both fp32-parity visibility-MLP engines: exact-fp32 FFMA and tcgen05 scaled-fp16 hi/lo x3."""
    from robir_b200 import ops
    old = ops.ENGINE["vis"]
    ops.ENGINE["vis"] = request.param
    yield request.param
    ops.ENGINE["vis"] = old


@pytest.mark.parametrize("terms,tol", [(3, 2e-6), (1, 2e-3)])
def test_tc_gemm_selftest(terms, tol):
    """tcgen05 machinery in isolation: TMEM-resident A, swizzled weight ring in streaming order, scaled fp16 operands.
    terms = 3: hi/lo 3-term split -- fp32-class accuracy (the bf16 split of round 1 sat at 3e-5 here);
    terms = 1: single-pass fp16 fast mode."""
    from robir_b200 import ops
    gen = torch.Generator().manual_seed(21)
    A = torch.randn(128, 256, generator=gen)
    W = torch.randn(256, 256, generator=gen) / 16
    D = ops.tc_selftest(A.cuda(), W.cuda(), terms).cpu()
    ref = (A.double() @ W.double().t()).float()
    err = (D - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, "tcgen05 GEMM self-test (terms=%d): max rel err %.3e" % (terms, err)


def test_sdf_network(golden, synth_sd16, model16):
    g = golden("nets")
    pts = g["pts"].cuda()
    out = model16.implicit_network(pts)
    ref = O.implicit_forward(synth_sd16, g["pts"])
    assert rel_err(out, ref) < REL
    assert rel_err(out[:, :8], g["sdf_feat_head"]) < REL
    grad = model16.implicit_network.gradient(pts)[:, 0]
    assert rel_err(grad, g["grad"]) < REL
    assert rel_err(model16.implicit_network.sdf(pts), ref[:, 0]) < REL
    # ragged / empty sizes
    for k in (0, 1, 15, 17, 63, 65):
        p = pts[:k]
        assert model16.implicit_network.sdf(p).shape == (k,) if k else True
        if k:
            s, gr = model16.implicit_network.sdf_and_normal(p)
            assert rel_err(s, ref[:k, 0]) < REL and rel_err(gr, g["grad"][:k]) < REL


def _vis_fn(sd):
    return lambda p, d: O.vis_network(sd, p, d)


@pytest.mark.parametrize("n", [37, 700])
def test_vis_mlp_backward_stagewise(synth_sd16, model16, engine, n):
    """Hot-kernel backward in isolation: d out / d sample_dir and d out / d weight against the oracle's autograd.
    n = 700 gives ~1400 tiles: every persistent CTA of the tensor-core engine walks ~10 tiles (cross-tile hand-offs)."""
    from robir_b200 import ops, sg_render
    sd = synth_sd16
    gen = torch.Generator().manual_seed(13)
    M, S = 16, 32
    pts = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1) * 0.33
    nrm = torch.nn.functional.normalize(pts + 0.1 * torch.randn(n, 3, generator=gen), dim=-1)
    dirs = torch.nn.functional.normalize(torch.randn(M * S, 3, generator=gen), dim=-1)
    w = torch.rand(M * S, generator=gen) + 0.1
    gup = torch.randn(n, M, generator=gen)
    d1, w1 = dirs.clone().requires_grad_(True), w.clone().requires_grad_(True)
    live = (nrm[:, None, :] * d1[None, :, :]).sum(-1) > 1e-6
    logits = O.vis_network(sd, pts[:, None, :].expand(n, M * S, 3)[live], d1[None].expand(n, M * S, 3)[live])
    vis = torch.zeros(n, M * S)
    vis[live] = torch.softmax(logits, -1)[:, 1]
    ref = (vis * w1[None]).reshape(n, M, S).sum(-1) / (w1.reshape(M, S).sum(-1)[None] + 1e-6)
    (ref * gup).sum().backward()
    d2, w2 = dirs.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    out = ops.diffuse_vis(pts.cuda(), nrm.cuda(), d2, w2, M, S, sg_render._weights_of(model16.visibility_network), True)
    assert rel_err(out, ref) < REL
    (out * gup.cuda()).sum().backward()
    grad_close(w2.grad, w1.grad, 1e-4, 1e-3, engine)
    grad_close(d2.grad, d1.grad, engine=engine)


def test_tc1_fast_mode_error_level(synth_sd16, model16):
    """ENGINE["vis"] = "tc1": single-pass fp16 on the same kernel (one MMA per product instead of three).  NOT the parity
    mode -- this test pins its error level: visibility within 1e-3 absolute (TF32-class; measured ~1e-4), direction
    gradient within 10 % relative L2 (borderline ReLU units flip at the 1e-4 level of its pre-activations)."""
    from robir_b200 import ops, sg_render
    sd = synth_sd16
    gen = torch.Generator().manual_seed(14)
    n, M, S = 300, 16, 32
    pts = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1) * 0.33
    nrm = torch.nn.functional.normalize(pts + 0.1 * torch.randn(n, 3, generator=gen), dim=-1)
    dirs = torch.nn.functional.normalize(torch.randn(M * S, 3, generator=gen), dim=-1)
    w = torch.rand(M * S, generator=gen) + 0.1
    gup = torch.randn(n, M, generator=gen)
    res = {}
    old = ops.ENGINE["vis"]
    try:
        for eng in ("tc", "tc1"):
            ops.ENGINE["vis"] = eng
            d2, w2 = dirs.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
            out = ops.diffuse_vis(pts.cuda(), nrm.cuda(), d2, w2, M, S,
                                  sg_render._weights_of(model16.visibility_network), True)
            (out * gup.cuda()).sum().backward()
            res[eng] = (out.detach().cpu(), d2.grad.cpu(), w2.grad.cpu())
    finally:
        ops.ENGINE["vis"] = old
    e_out = (res["tc1"][0] - res["tc"][0]).abs().max().item()
    e_gd = ((res["tc1"][1] - res["tc"][1]).norm() / res["tc"][1].norm()).item()
    e_gw = ((res["tc1"][2] - res["tc"][2]).norm() / res["tc"][2].norm()).item()
    print("\ntc1 (single-pass fp16) vs tc (fp32 parity): light_vis max abs %.2e, d/d dir rel L2 %.2e, d/d w rel L2 %.2e"
          % (e_out, e_gd, e_gw))
    assert 0 < e_out < 1e-3 and e_gd < 0.1 and e_gw < 1e-2


def test_diffuse_visibility_fwd_bwd(synth_sd16, model16, engine):
    from robir_b200 import rng, sg_render
    sd = synth_sd16
    gen = torch.Generator().manual_seed(3)
    n, M, S = 70, 16, 32
    pts = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1) * 0.33
    nrm = torch.nn.functional.normalize(pts + 0.1 * torch.randn(n, 3, generator=gen), dim=-1)
    lobes = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen), dim=-1)
    lam = torch.rand(M, 1, generator=gen) * 60 + 0.5
    ut, up = torch.rand(M, S, generator=gen), torch.rand(M, S, generator=gen)
    gup = torch.randn(M, n, generator=gen)
    lo, la = lobes.clone().requires_grad_(True), lam.clone().requires_grad_(True)
    ref = O.get_diffuse_visibility(pts, nrm, _vis_fn(sd), lo, la, ut, up)
    (ref * gup).sum().backward()
    lo2, la2 = lobes.cuda().requires_grad_(True), lam.cuda().requires_grad_(True)
    with rng.replay([ut, up]):
        out = sg_render.get_diffuse_visibility(pts.cuda(), nrm.cuda(), model16.visibility_network, lo2, la2, nsamp=S)
    (out * gup.cuda()).sum().backward()
    assert out.shape == ref.shape == (M, n)
    assert rel_err(out, ref) < REL
    grad_close(lo2.grad, lo.grad, engine=engine)
    grad_close(la2.grad, la.grad, engine=engine)
    # testing mode (no_grad VisModel) gives the same values
    with rng.replay([ut, up]), torch.no_grad():
        out_t = sg_render.get_diffuse_visibility(pts.cuda(), nrm.cuda(), model16.visibility_network, lobes.cuda(),
                                                 lam.cuda(), nsamp=S, testing=True)
    assert rel_err(out_t, ref) < REL


def test_specular_visibility_fwd_bwd(synth_sd16, model16, engine):
    from robir_b200 import rng, sg_render
    sd = synth_sd16
    gen = torch.Generator().manual_seed(4)
    n, S = 77, 8
    pts = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1) * 0.33
    nrm = torch.nn.functional.normalize(pts, dim=-1)
    view = torch.nn.functional.normalize(nrm + 0.7 * torch.randn(n, 3, generator=gen), dim=-1)
    rough = torch.rand(n, 1, generator=gen) * 0.9 + 0.09
    ut, up = torch.rand(n, S, generator=gen), torch.rand(n, S, generator=gen)
    gup = torch.randn(n, generator=gen)
    for inv in (False, True):
        r1 = rough.clone().requires_grad_(True)
        wl, wlam = sg_render._spec_warp(nrm, view, r1)
        ref = O.get_specular_visibility(pts, nrm, view, _vis_fn(sd), wl, wlam, ut, up, inv=inv)
        (ref * gup).sum().backward()
        r2 = rough.cuda().requires_grad_(True)
        wl2, wlam2 = sg_render._spec_warp(nrm.cuda(), view.cuda(), r2)
        with rng.replay([ut, up]):
            out = sg_render.get_specular_visibility(pts.cuda(), nrm.cuda(), view.cuda(), model16.visibility_network,
                                                    wl2, wlam2, nsamp=S, inv=inv)
        (out * gup.cuda()).sum().backward()
        assert rel_err(out, ref) < REL
        grad_close(r2.grad, r1.grad, engine=engine)


def test_octree_cast_on_oracle_tree(golden, oracle_octrees):
    """Same arrays, CUDA walk vs. oracle walk: masks identical, distances bit-exact (no FMA contraction on device)."""
    from robir_b200 import ops
    g = golden("octree")
    prim, sec = oracle_octrees
    a = prim.arrays()
    tree = ops.PackedOctree(a["root"], a["boxes"], a["non_leaf"], a["links"], a["grid"], a["sdf_val"], a["sdf_grad"],
                            a["min_step"], "cuda")
    N = g["ray_dirs"].shape[1]
    x, hit, t, cnt = ops.octree_cast(tree, g["cam_loc"].cuda(), g["ray_dirs"].reshape(-1, 3).cuda(), max_iter=-1,
                                     o_div=N, return_stats=True)
    xo, ho, to = prim.trace(g["cam_loc"], g["ray_dirs"])
    assert torch.equal(hit.cpu(), ho) and torch.equal(hit.cpu(), g["prim_mask"])
    assert torch.equal(t.cpu()[ho], to[ho]) and torch.equal(x.cpu()[ho], xo[ho])
    assert (t.cpu()[ho] - g["prim_t"][ho]).abs().max() < 2e-5
    assert int(cnt[-6]) > 10  # lock-step iterations executed
    S = g["sec_d"].shape[1]
    x, hit, t = ops.octree_cast(tree, g["sec_o"].cuda(), g["sec_d"].reshape(-1, 3).cuda(), max_iter=32, o_div=S)
    xo, ho, to = sec.trace(g["sec_o"], g["sec_d"])
    assert torch.equal(hit.cpu(), ho) and torch.equal(t.cpu(), to) and torch.equal(x.cpu(), xo)
    x, hit, t = ops.octree_cast(tree, g["edge_o"].cuda(), g["edge_d"].reshape(-1, 3).cuda(), max_iter=-1, o_div=1)
    assert hit.tolist() == [False, True] and torch.isnan(t[0]) and abs(float(t[1]) - float(g["edge_t"][1])) < 2e-5
    # empty call
    x, hit, t = ops.octree_cast(tree, torch.zeros(0, 3).cuda(), torch.zeros(0, 3).cuda())
    assert x.shape == (0, 3) and hit.shape == (0,)


def test_octree_build_on_gpu(golden, model16, oracle_octrees):
    g = golden("octree")
    model16.generate()
    tree = model16.ray_tracer.sdf_octree
    prim, _ = oracle_octrees
    # split decisions compare |sdf| with a threshold; fp32 noise may flip a handful of nodes out of ~10^6
    assert abs(tree.n_nodes - int(g["fp_n_nodes"])) <= 64
    if tree.n_nodes == prim.boxes.shape[0]:
        assert (tree.nodes[:, 7].cpu() - prim.sdf_val).abs().max() < 5e-6
        assert (tree.sdf_grad.cpu() - prim.sdf_grad).abs().max() < 5e-5


@pytest.mark.parametrize("normal_grad", [False, True])
def test_sg_render_kernel(synth_sd16, normal_grad):
    """normal_grad: the instantiation that also differentiates with respect to the shading normal (CESR)."""
    from robir_b200 import ops
    from test_host_math import _oracle_sg, _rand_scene
    n, M, Mi = 50, 16, 24
    s = _rand_scene(n, M, Mi, seed=9)
    names = ["rough", "albedo", "spec", "lgt", "ind", "lv", "bvd", "bvi", "integ"] + (["normal"] if normal_grad else [])
    out, leaves = _oracle_sg(s, names)
    keys = ["sg_rgb", "sg_specular_rgb", "sg_diffuse_rgb", "vis_shadow", "indir_rgb", "indir_specular_rgb",
            "indir_diffuse_rgb"]
    gen = torch.Generator().manual_seed(5)
    gup = [torch.randn(n, 3, generator=gen) for _ in keys]
    gup[3].zero_()
    sum((out[k] * g).sum() for k, g in zip(keys, gup)).backward()
    c = {k: s[k].cuda().requires_grad_(k in names) for k in s}
    res = ops.sg_render(c["normal"], c["view"], c["rough"], c["albedo"], c["spec"].abs().reshape(1), c["lgt"], c["ind"],
                        c["lv"], c["bvd"], c["bvi"], c["integ"], False)
    for k, r in zip(keys, res):
        assert rel_err(r, out[k]) < REL, k
    sum((r * g.cuda()).sum() for r, g in zip(res, gup) if r.requires_grad).backward()
    for k in names:
        ref = leaves[k].grad
        got = c[k].grad.cpu()
        # roughness enters as 2/r^4 (up to 3e4 at r = 0.09): ill-conditioned, looser bound
        tol = 3e-3 if k == "rough" else 5e-4
        assert torch.isfinite(got).all(), k
        assert (got - ref).abs().max().item() <= tol * max(1e-6, ref.abs().max().item()), k


def test_pbr_step_vs_golden(golden, synth_sd16, model16, engine):
    """Full IDRNetwork.forward('Material') + loss + backward against the reference's golden outputs and gradients."""
    from robir_b200 import rng
    from robir_b200.loss import InvLoss, pbr_step_loss
    g = golden("pbr_step")
    model16.generate()
    model16.zero_grad()
    N = g["pix"].shape[0]
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(g["pix"]).items()}
    inp["hdr_shift"] = model16.gamma.hdr_shift.as_input().expand(N, 1)
    with rng.replay([g["rnd_%d" % i] for i in range(9)]):
        out = model16(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
    mask = out["network_object_mask"].cpu()
    assert (mask != g["out_network_object_mask"]).sum() == 0, "tracer mask differs from the reference"
    for k in [k[4:] for k in g if k.startswith("out_") and k != "out_network_object_mask"]:
        assert out[k].shape == g["out_" + k].shape, k
        assert rel_err(out[k], g["out_" + k]) < REL, (k, rel_err(out[k], g["out_" + k]))
    loss, _ = pbr_step_loss(model16, InvLoss(), out, {"rgb": g["gt"]})
    assert abs(loss.item() - g["loss"].item()) < 1e-4
    loss.backward()
    mat = model16.envmap_material_network
    dec, enc = mat.spec_brdf_encoder_layer.brdf_decoder_layer, mat.spec_brdf_encoder_layer.brdf_encoder_layer
    checks = [(mat.lgtSGs.grad, g["g_lgtSGs"]), (mat.specular_reflectance.grad, g["g_spec"]),
              (model16.gamma.hdr_shift.adapt_illum.grad, g["g_adapt"]), (dec[4].bias.grad, g["g_dec4_bias"]),
              (dec[4].weight.grad, g["g_dec4_weight"]), (enc[0].bias.grad, g["g_enc0_bias"]),
              (enc[8].weight.grad.sum(0), g["g_enc8_weight_sum"])]
    for a, b in checks:
        grad_close(a, b, 5e-3, 3e-2, engine)


def test_fused_loss_matches_torch_glue(golden, model16):
    """csrc/loss.cu (value + all gradients in one launch) against the elementwise torch restatement of
    model/loss.py:61-125 + train_pbr.py:313-346 on the same forward graph, L1 and L2."""
    from robir_b200 import loss as L, rng
    g = golden("pbr_step")
    model16.generate()
    N = g["pix"].shape[0]
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(g["pix"]).items()}
    inp["hdr_shift"] = model16.gamma.hdr_shift.as_input().expand(N, 1)
    with rng.replay([g["rnd_%d" % i] for i in range(9)]):
        out = model16(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
    mat = model16.envmap_material_network
    enc = mat.spec_brdf_encoder_layer.brdf_encoder_layer
    leaves = [mat.lgtSGs, mat.specular_reflectance, model16.gamma.hdr_shift.adapt_illum, enc[0].bias, enc[8].weight]
    gen = torch.Generator().manual_seed(5)
    gt = {"rgb": torch.rand(1, N, 3, generator=gen)}
    for loss_type in ("L1", "L2"):
        fn = L.InvLoss(loss_type=loss_type)
        L.FUSED_LOSS = False
        try:
            ref, ref_parts = L.pbr_step_loss(model16, fn, out, gt)
        finally:
            L.FUSED_LOSS = True
        got, parts = L.pbr_step_loss(model16, fn, out, gt)
        assert abs(got.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item())), loss_type
        for k in ("sg_rgb_loss", "kl_loss", "latent_smooth_loss", "loss"):
            assert abs(parts[k].item() - ref_parts[k].item()) < 1e-5 * max(1.0, abs(ref_parts[k].item())), k
        g_ref = torch.autograd.grad(ref, leaves, retain_graph=True)
        g_got = torch.autograd.grad(got, leaves, retain_graph=True)
        for a, b in zip(g_got, g_ref):
            assert (a - b).abs().max().item() <= 2e-4 * max(1e-7, b.abs().max().item()), loss_type


def test_pbr_forward_properties_full_size(model16, synth_sd16):
    """BASELINE-size batch (1024 rays): size-independent properties -- determinism under replayed randoms, outputs of
    non-hit rays keep the reference's 1.0 fill, and hit-ray outputs do not depend on which other rays share the batch
    once the batch-coupled scalars (octree live count, sg_range) are pinned by using the same hit set."""
    from robir_b200 import rng
    model16.generate()
    N = 1024
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(11, n=N, crop=500)).items()}
    inp["hdr_shift"] = torch.full((N, 1), 0.5).cuda()
    with rng.record() as tape, torch.no_grad():
        a = model16(inp, trainstage="Material", train_spec=True)
    with rng.replay(tape), torch.no_grad():
        b = model16(inp, trainstage="Material", train_spec=True)
    m = a["network_object_mask"]
    assert 0 < int(m.sum()) < N
    for k in ("sg_rgb", "indir_rgb", "vis_shadow", "roughness"):
        assert torch.equal(a[k], b[k]), k
        assert torch.all(a[k][~m] == 1.0), k
        assert torch.isfinite(a[k]).all(), k
    assert (a["sg_rgb"][m] >= 0).all()


def test_static_shape_mode_matches_compacted_mode(model16):
    """static_shapes=True (fixed-capacity batch, device-side hit compaction; what the CUDA-graph step uses) vs. the
    reference-shaped path."""
    from robir_b200 import rng
    model16.generate()
    N = 512
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(12, n=N, crop=460)).items()}
    inp["hdr_shift"] = torch.full((N, 1), 0.5).cuda()
    with rng.record() as tape, torch.no_grad():
        a = model16(inp, trainstage="Material", train_spec=True)
    m = a["network_object_mask"].cpu()
    assert 0 < int(m.sum()) < N
    tape_static = []
    for t in tape:
        if t.shape[0] == int(m.sum()) and t.shape[0] != 16:      # per-hit draws -> rows [0, n_hit) of [N, .] (the
            full = torch.zeros(N, *t.shape[1:])                  # static path compacts the hit rays to the front)
            full[:t.shape[0]] = t
            tape_static.append(full)
        else:
            tape_static.append(t)
    model16.static_shapes = True
    try:
        with rng.replay(tape_static), torch.no_grad():
            b = model16(inp, trainstage="Material", train_spec=True)
    finally:
        model16.static_shapes = False
    assert set(a.keys()) == set(b.keys())
    for k in a:
        if a[k].dtype == torch.bool:
            assert torch.equal(a[k], b[k]), k
        elif k != "points":
            assert a[k].shape == b[k].shape, k
            # the two modes run different launch shapes (atomics, row counts): a visibility sample within an ulp of the
            # hard culling predicate cos(normal, direction) > 0 may land on either side and moves ONE ray by ~1/32 of a
            # lobe weight; everything else agrees to the fp32 floor
            scale = max(1.0, float(a[k].abs().max()))
            d = (b[k].float() - a[k].float()).abs().reshape(a[k].shape[0], -1).amax(1) / scale if a[k].dim() > 0 and \
                a[k].shape[0] == N else (b[k].float() - a[k].float()).abs().reshape(1, -1).amax(1) / scale
            bad = d > 1e-5
            assert int(bad.sum()) <= 2 and float(d.max()) < 5e-3, (k, int(bad.sum()), float(d.max()))


def test_strong_sharding_replicated_walk_is_exact(model16):
    """dist.STRONG_SHARDING with input["shard"]: every rank walks the WHOLE batch through the octree (the walk's sample
    count per lock-step iteration depends on the live rays of the whole batch, utils/octree.py:542-548) and shades its
    slice -- the traced quantities of the slices are bit-identical to the single-rank forward of the full batch, which a
    rank tracing only its own rays does not guarantee.  (The loss / gradient side of strong sharding is covered by the
    2-rank gloo tests in tests/test_dist_cpu.py.)"""
    from robir_b200 import dist as rdist, rng
    model16.generate()
    N = 600
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(21, n=N, crop=520)).items()}
    inp["hdr_shift"] = torch.full((N, 1), 0.5).cuda()
    torch.manual_seed(5)
    with torch.no_grad():
        full = model16(inp, trainstage="Material", train_spec=True)
    parts = []
    for rank in range(3):
        lo, hi = rdist.shard_rays(N, rank, 3)
        torch.manual_seed(5 + rank)
        with torch.no_grad():
            parts.append(model16(dict(inp, shard=(lo, hi)), trainstage="Material", train_spec=True))
        assert parts[-1]["network_object_mask"].shape[0] == hi - lo
    for k in ("network_object_mask", "object_mask", "points", "ray_dirs", "sdf_output", "normals", "diffuse_albedo"):
        cat = torch.cat([p[k] for p in parts], 0)
        assert cat.shape == full[k].shape, k
        if k in ("sdf_output", "normals", "diffuse_albedo"):        # per-point networks: batch size picks the engine
            assert rel_err(cat, full[k]) < REL, k
        else:
            assert torch.equal(cat, full[k]), k
    assert 0 < int(full["network_object_mask"].sum()) < N


def test_graphed_step_runs_and_trains(synth_sd16):
    """CUDA-graph capture of the whole PBR training step: the loss goes down, and the captured graph re-packs the
    trained weights on every replay whatever optimizer implementation updates them (torch's fused Adam does not bump
    tensor versions: the trajectory must equal the one of the default implementation)."""
    import robir_b200
    from robir_b200 import graph, rng
    from robir_b200.loss import InvLoss
    rng.set_mode("device")
    try:
        finals = {}
        for fused in (False, True):
            torch.manual_seed(7)
            torch.cuda.manual_seed(7)
            m = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=16)))
            m.load_state_dict(synth_sd16, strict=True)
            m.cuda().train()
            m.generate()
            params = list(m.gamma.parameters()) + list(m.envmap_material_network.parameters())
            opt = torch.optim.Adam(params, lr=5e-4, capturable=True, fused=fused)
            N = 256
            step = graph.GraphedPBRStep(m, InvLoss(), opt, N, synthetic.camera_pose().cuda(),
                                        synthetic.camera_intrinsics().cuda())
            inp = synthetic.camera_inputs(synthetic.training_pixels(3, n=N, crop=400))
            gt = torch.full((1, N, 3), 0.3).cuda()
            losses = [float(step(inp["uv"].cuda(), inp["object_mask"].cuda(), gt)) for _ in range(25)]
            assert all(np.isfinite(losses)) and losses[-1] < 0.8 * losses[0], losses
            assert step.launches_per_step > 20
            finals[fused] = losses[-1]
        assert abs(finals[True] - finals[False]) < 0.05 * abs(finals[False]), finals
    finally:
        rng.set_mode("cpu")


def test_fused_small_networks_vs_golden_and_library(golden, model16):
    """Rows a6/a7: fused MLP kernels (csrc/mlp.cu) vs the reference goldens (forward) and vs the plain library-GEMM
    evaluation of the same modules on the GPU (forward + all gradients)."""
    from robir_b200 import networks, rng
    g = golden("nets")
    pts, hs = g["pts"].cuda(), g["hdr_shift"].cuda()
    with rng.replay([g["noise_indir"]]):
        sgs, env = model16.indirect_illum_network(pts, hs)
    assert rel_err(sgs, g["indir_sgs"]) < REL and rel_err(env, g["indir_env"]) < REL
    with rng.replay([g["noise_brdf"], g["noise_nrm"]]):
        mat = model16.envmap_material_network(pts, train_spec=True)
    for a, b in [("sg_roughness", "roughness"), ("sg_diffuse_albedo", "albedo"), ("sg_metallic", "metallic"),
                 ("sg_normal_map", "normal_map"), ("random_xi_roughness", "xi_roughness"),
                 ("random_xi_diffuse_albedo", "xi_albedo")]:
        assert rel_err(mat[a], g[b]) < REL, a

    def run(fused):
        networks.FUSED_MLP = fused
        model16.zero_grad()
        h = hs.clone().requires_grad_(True)
        model16.indirect_illum_network.train_weights = True
        try:
            with rng.replay([g["noise_indir"], g["noise_brdf"], g["noise_nrm"]]):
                s, e = model16.indirect_illum_network(pts, h)
                m = model16.envmap_material_network(pts, train_spec=True)
            gen = torch.Generator().manual_seed(8)
            w = [torch.randn(t.shape, generator=gen).cuda() for t in (s, e, m["sg_roughness"], m["sg_diffuse_albedo"],
                                                                      m["random_xi_roughness"])]
            loss = sum((t * wi).sum() for t, wi in zip((s, e, m["sg_roughness"], m["sg_diffuse_albedo"],
                                                         m["random_xi_roughness"]), w))
            loss.backward()
        finally:
            networks.FUSED_MLP = True
            model16.indirect_illum_network.train_weights = False
        grads = {k: p.grad.clone() for k, p in model16.named_parameters() if p.grad is not None}
        return float(loss), h.grad.clone(), grads

    l1, gh1, g1 = run(True)
    l0, gh0, g0 = run(False)
    assert abs(l1 - l0) < 1e-4 * max(1.0, abs(l0))
    grad_close(gh1, gh0, 1e-4, 1e-3)
    assert set(g1) == set(g0) and len(g0) >= 40
    for k in g0:
        grad_close(g1[k], g0[k], 2e-4, 2e-3)


def _with_oracle_tree(model, prim):
    """Give both tracers of the model the oracle's octree arrays (so that hit masks are comparable bit for bit)."""
    from robir_b200 import ops
    a = prim.arrays()
    tree = ops.PackedOctree(a["root"], a["boxes"], a["non_leaf"], a["links"], a["grid"], a["sdf_val"], a["sdf_grad"],
                            a["min_step"], "cuda")
    model.ray_tracer.sdf_octree = tree
    model.octree_ray_tracer.sdf_octree = tree
    return tree


def test_borrow_color_vs_golden(golden, model16):
    """Row a13: batch_borrow_color = fused SDF value/normal/feature kernel + fused colour MLP + 16-sample NeuS render."""
    g = golden("nets")
    col = model16.implicit_network.batch_borrow_color(g["pts"].cuda(), g["vdirs"].cuda())
    assert rel_err(col, g["borrow_color"]) < REL
    assert model16.implicit_network.batch_borrow_color(g["pts"][:0].cuda(), g["vdirs"][:0].cuda()).shape == (0, 3)
    logits = model16.visibility_network(g["pts"].cuda(), g["vdirs"].cuda())
    assert rel_err(logits, g["vis_logits"]) < REL


def test_vis_stage_vs_golden(golden, synth_sd16, model16, oracle_octrees):
    """Row a13 / config 3: forward('Illum') + trace_radiance against the reference's golden outputs, then the Vis-stage
    losses (model/loss.py:144-179) and their gradients w.r.t. the visibility / indirect networks against the oracle."""
    from robir_b200 import rng
    from robir_b200.loss import IllumLoss
    g, sd = golden("vis_stage"), synth_sd16
    prim, sec = oracle_octrees
    _with_oracle_tree(model16, prim)
    model16.zero_grad()
    model16.indirect_illum_network.train_weights = True
    try:
        inp = {k: v.cuda() for k, v in synthetic.camera_inputs(g["pix"]).items()}
        inp["hdr_shift"] = g["rnd_0"].cuda()
        with rng.replay([g["rnd_1"], g["rnd_2"]]):
            out = model16(inp, trainstage="Illum")
        assert torch.equal(out["network_object_mask"].cpu(), g["mask"])
        for k in ["indirect_sgs", "indir_integral", "normals", "points"]:
            assert rel_err(out[k], g[k]) < REL, k
        with rng.replay([g["rnd_3"], g["rnd_4"]]):
            tr = model16.trace_radiance(out, nsamp=16)
        for k in ["gt_vis", "indir_mask"]:
            assert torch.equal(tr[k].cpu(), g["tr_" + k]), k
        for k in ["trace_radiance", "sample_dirs", "pred_vis", "gt_integral"]:
            assert tr[k].shape == g["tr_" + k].shape, k
            assert rel_err(tr[k], g["tr_" + k]) < REL, (k, rel_err(tr[k], g["tr_" + k]))
        rad_loss, vis_loss = IllumLoss()(out, tr, 0.0)
        (rad_loss + vis_loss).backward()
    finally:
        model16.indirect_illum_network.train_weights = False
    # ---- oracle: same losses by autograd over the state dict
    sdg = {k: v.clone() for k, v in sd.items()}
    train = [k for k in sdg if k.startswith("visibility_network.") or k.startswith("indirect_illum_network.")]
    for k in train:
        sdg[k].requires_grad_(True)
    inp_o = synthetic.camera_inputs(g["pix"])
    inp_o["hdr_shift"] = g["rnd_0"]
    out_o = P.idr_forward(sdg, inp_o, lambda c, m, d: prim.trace(c, d),
                          dict(indir_noise=g["rnd_1"], normal_noise=g["rnd_2"]), trainstage="Illum")
    tr_o = P.trace_radiance(sdg, out_o, lambda c, m, d: sec.trace(c, d), g["rnd_3"], g["rnd_4"], 16)
    rad_o, vis_o = O.illum_loss(out_o, tr_o, 0.0)
    (rad_o + vis_o).backward()
    assert abs(float(rad_loss) - float(rad_o)) < 1e-4 * max(1.0, abs(float(rad_o)))
    assert abs(float(vis_loss) - float(vis_o)) < 1e-4
    got = dict(model16.named_parameters())
    n_checked = 0
    for k in train:
        if sdg[k].grad is None:
            continue
        assert got[k].grad is not None, k
        # the 512-wide chains run on the tensor-core layer engine (bf16 hi/lo operands): see grad_close
        grad_close(got[k].grad, sdg[k].grad, 2e-3, 2e-2, engine=ops_engine_mlp())
        n_checked += 1
    assert n_checked >= 20


def test_vis_stage_full_size_properties(model16):
    """Config-3 sized call (N=256 primary rays, nsamp=512 -> 131 072 secondary rays): shapes, masks and invariants that
    do not need the oracle: back-facing samples carry no radiance, gt_integral is the cosine-weighted hemisphere mean,
    rows of non-hit primary rays stay zero, replayed randoms reproduce the result bit for bit."""
    from robir_b200 import rng
    model16.generate()
    N, S = 256, 512
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(5, n=N, crop=420)).items()}
    inp["hdr_shift"] = torch.rand(N, 1, generator=torch.Generator().manual_seed(3)).cuda()
    with torch.no_grad():
        with rng.record() as tape:
            out = model16(inp, trainstage="Illum")
            tr = model16.trace_radiance(out, nsamp=S)
        with rng.replay(tape):
            out2 = model16(inp, trainstage="Illum")
            tr2 = model16.trace_radiance(out2, nsamp=S)
    m = out["network_object_mask"]
    n = int(m.sum())
    assert 0 < n < N
    assert tr["trace_radiance"].shape == (N, S, 3) and tr["sample_dirs"].shape == (n, S, 3)
    assert tr["gt_vis"].shape == (N, S, 1) and tr["pred_vis"].shape == (N, S, 2) and tr["indir_mask"].shape == (N, S)
    for k in tr:
        assert torch.equal(tr[k], tr2[k]), k
        assert torch.isfinite(tr[k].float()).all(), k
    assert not tr["gt_vis"][~m].any() and (tr["trace_radiance"][~m] == 0).all() and (tr["pred_vis"][~m] == 0).all()
    nrm = out["normals"][m]
    nrm = nrm / nrm.norm(dim=-1, keepdim=True).clamp_min(1e-4)
    cosv = (nrm[:, None, :] * tr["sample_dirs"]).sum(-1)
    assert (tr["trace_radiance"][m][cosv < 0] == 0).all()
    assert not tr["indir_mask"][m][cosv < 0].any()
    want = (tr["trace_radiance"][m] * torch.relu(cosv)[..., None]).sum(1) / (cosv >= 0).sum(-1, keepdim=True).clamp_min(1e-4)
    assert rel_err(tr["gt_integral"][m], want) < 1e-5
    assert (tr["sample_dirs"].norm(dim=-1) - 1).abs().max() < 1e-5


def _trace_agreement(got, ref, tol=2e-4, frac=0.98):
    """Sphere-tracer outputs vs the oracle: the march makes threshold decisions on fp32 SDF values (5e-5 convergence
    test, sign of samples), so a handful of borderline rays may take another branch on the GPU; the rest must agree to
    fp32 accuracy.  -> fraction of rays with identical mask and |dt|, |dp| < tol."""
    (p, m, t), (p2, m2, t2) = got, ref
    p, m, t = p.cpu(), m.cpu(), t.cpu()
    same = (m == m2) & ((t - t2).abs() < tol) & ((p - p2).abs().max(-1)[0] < tol)
    agree = same.float().mean().item()
    assert agree >= frac, "sphere tracer agrees with the oracle on %.1f %% of the rays only" % (100 * agree)
    return agree


def test_camera_rays_skew_and_quaternion_pose():
    """Row a2, directly: robir_camera_rays against the oracle's get_camera_params + lift (utils/rend_util.py:51-97, pinned to
    the reference in tests/test_oracle_vs_reference.py) with non-zero skew, an off-centre principal point, DTU-sized
    intrinsics, a 4x4 pose and the 7-vector (quaternion + position) pose; ragged and empty batches."""
    from robir_b200 import ops
    gen = torch.Generator().manual_seed(77)
    uv = torch.rand(1, 300, 2, generator=gen) * torch.tensor([1600.0, 1200.0])
    K = torch.eye(3)[None].clone()
    K[0, 0, 0], K[0, 1, 1], K[0, 0, 2], K[0, 1, 2], K[0, 0, 1] = 2892.0, 2883.0, 823.2, 619.1, 7.5
    pose7 = torch.tensor([[0.31, -0.62, 0.48, 0.53, 1.3, -0.4, 2.2]])
    for pose in (pose7, synthetic.camera_pose()):
        rd_o, cl_o = O.camera_rays(uv, pose, K)
        for n in (300, 1, 0, 37):
            rd, cl = ops.camera_rays(uv[:, :n].cuda(), pose.cuda(), K.cuda())
            assert rd.shape == (1, n, 3) and cl.shape == (1, 3)
            assert (cl.cpu() - cl_o).abs().max().item() < 1e-6
            if n:
                assert (rd.cpu() - rd_o[:, :n]).abs().max().item() < 1e-6
                assert (rd.norm(dim=-1) - 1).abs().max().item() < 1e-6
    with pytest.raises(Exception):
        ops.camera_rays(uv.cuda(), torch.zeros(1, 5).cuda(), K.cuda())


def test_sphere_tracer_vs_golden(golden, synth_sd16, model16):
    """Row a3: persistent march kernels (csrc/sphere_trace.cu) vs the reference's RayTracing outputs (eval and training
    mode, n_steps = 32, hotdog.conf tracer settings)."""
    from robir_b200 import ops, rng
    from robir_b200.sphere_tracing import RayTracing
    g = golden("raytracing")
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(g["pix"]).items()}
    rd, cl = ops.camera_rays(inp["uv"], inp["pose"], inp["intrinsics"])
    tracer = RayTracing(line_step_iters=3, n_steps=32, n_rootfind_steps=32).bind(model16.implicit_network).cuda()
    om = g["object_mask"].cuda()
    for tag, training in (("eval", False), ("train", True)):
        tracer.train(training)
        with rng.replay([g["uniform"]] if training else []):
            got = tracer(sdf=model16.implicit_network.sdf, cam_loc=cl, object_mask=om, ray_directions=rd)
        assert got[0].shape == (256, 3) and got[1].dtype == torch.bool and got[2].shape == (256,)
        _trace_agreement(got, (g[tag + "_points"], g[tag + "_mask"], g[tag + "_t"]))
        assert int(tracer.last_counters[4]) > 256      # SDF queries executed inline
    # ragged / empty batches
    for k in (0, 1, 17):
        p, m, t = tracer(sdf=model16.implicit_network.sdf, cam_loc=cl, object_mask=om[:k], ray_directions=rd[:, :k])
        assert p.shape == (k, 3) and m.shape == (k,) and t.shape == (k,)


def test_sphere_tracer_vs_oracle_perturbed_and_sampler(synth_sd16):
    """Row a3 on a perturbed (non-spherical) SDF, 1024 rays, n_steps = 128 (BASELINE config 2), with a tight
    sphere-tracing budget so that the sampler + secant phases carry most rays; per-origin batches (o_div = 1)."""
    import robir_b200
    from robir_b200 import ops
    from robir_b200.sphere_tracing import RayTracing
    sd = synthetic.synthetic_state_dict(0, num_lgt_sgs=16, perturb=0.05)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=16)))
    model.load_state_dict(sd, strict=True)
    model.cuda()
    sdf_o = lambda x: O.implicit_forward(sd, x)[:, 0]
    inp = synthetic.camera_inputs(synthetic.training_pixels(4, n=1024, crop=460))
    rd_o, cl_o = O.camera_rays(inp["uv"], inp["pose"], inp["intrinsics"])
    om = torch.rand(1024, generator=torch.Generator().manual_seed(2)) > 0.2
    uni = torch.rand(128, generator=torch.Generator().manual_seed(3))
    for iters, training in ((10, False), (2, False), (2, True)):
        with torch.no_grad():
            ref = T.ray_tracing(sdf_o, cl_o, om, rd_o, line_step_iters=3, sphere_tracing_iters=iters, n_steps=128,
                                n_secant_steps=32, training=training, uniform_steps=uni)
        got = ops.sphere_trace(model.implicit_network._w, cl_o.cuda(), rd_o.cuda(), om.cuda(), line_step_iters=3,
                               sphere_tracing_iters=iters, n_steps=128, n_secant_steps=32, training=training,
                               uniform_steps=uni.cuda(), return_stats=True)
        _trace_agreement(got[:3], ref, frac=0.97)
        cnt = got[3].cpu()
        if iters == 2:
            assert int(cnt[0]) > 100 and int(cnt[1]) > 50     # sampler and secant rays
    # secondary-ray style call: one origin per ray
    o = torch.nn.functional.normalize(torch.randn(300, 3, generator=torch.Generator().manual_seed(5)), dim=-1) * 0.9
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(300, 3, generator=torch.Generator().manual_seed(6)), dim=-1)
    with torch.no_grad():
        ref = T.ray_tracing(sdf_o, o, torch.ones(300, dtype=torch.bool), d[:, None, :], line_step_iters=3, n_steps=64,
                            n_secant_steps=8)
    got = ops.sphere_trace(model.implicit_network._w, o.cuda(), d[:, None, :].cuda(), None, line_step_iters=3,
                           n_steps=64, n_secant_steps=8)
    _trace_agreement(got, ref, frac=0.97)


def test_idr_network_with_sphere_tracer(synth_sd16):
    """use_octree=False model: forward('Material') runs on the sphere tracer and agrees with the octree model on the
    rays both tracers hit (two different surface finders on the same SDF: points agree to ~1e-3)."""
    import robir_b200
    from robir_b200 import rng
    conf = dict(envmap_material_network=dict(num_lgt_sgs=16), use_octree=False,
                ray_tracer=dict(line_step_iters=3, n_steps=128, n_rootfind_steps=32))
    model = robir_b200.IDRNetwork(conf)
    model.load_state_dict(synth_sd16, strict=True)
    model.cuda().eval()
    model.generate()
    N = 256
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(7, n=N, crop=420)).items()}
    inp["hdr_shift"] = torch.full((N, 1), 0.5).cuda()
    with torch.no_grad():
        out = model(inp, trainstage="Material", train_spec=True)
    m = out["network_object_mask"]
    assert 0 < int(m.sum()) < N
    sdf = model.implicit_network.sdf(out["points"][m])
    assert sdf.abs().max() < 1e-3                       # hit points lie on the zero level set
    for k in ("sg_rgb", "indir_rgb", "normals"):
        assert torch.isfinite(out[k]).all() and (out[k][~m] == 1).all(), k


def test_wgrad_kernel_shapes_and_active_rows():
    """csrc/mlp.cu wgrad_kernel against torch matmul: ragged shapes, row splits, segments with an inactive tail."""
    import ctypes
    from robir_b200._lib import check, lib, ptr, stream
    gen = torch.Generator().manual_seed(11)
    for n, N, K, ldg, lda, seg, n_act, splits in [(77, 5, 63, 256, 64, 0, None, 1), (1024, 512, 512, 512, 512, 0, None, 2),
                                                  (2048, 32, 512, 256, 512, 1024, 300, 16), (96, 144, 512, 256, 512, 0, 40, 4),
                                                  (1, 3, 7, 8, 8, 0, None, 1)]:
        G = torch.randn(n, ldg, generator=gen).cuda()
        A = torch.randn(n, lda, generator=gen).cuda()
        Gr, Ar = G.clone(), A.clone()
        na = None
        if n_act is not None:
            sg = seg or n
            rows = torch.arange(n) % sg
            # rows outside the active head carry zero gradient in the product; whole inactive 32-row chunks are skipped
            Gr[(rows >= n_act).cuda()] = 0
            G = Gr.clone()
            na = torch.tensor([n_act], dtype=torch.int32).cuda()
        tiles = ((N + 63) // 64) * ((K + 63) // 64)
        part = torch.empty(splits * tiles * 4160).cuda()
        tick = torch.zeros(tiles, dtype=torch.int32).cuda()
        dW, db = torch.empty(N, K).cuda(), torch.empty(N).cuda()
        for _ in range(2):        # twice: the ticket counters must come back to zero
            check(lib().robir_mlp_wgrad(ptr(G), ldg, ptr(A), lda, n, N, K, ptr(na), seg, splits,
                                        ptr(part) if splits > 1 else None, ptr(tick), ptr(dW), ptr(db), stream()))
        ref_w = (Gr[:, :N].double().t() @ Ar[:, :K].double()).float()
        ref_b = Gr[:, :N].double().sum(0).float()
        assert rel_err(dW, ref_w) < 1e-5 and rel_err(db, ref_b) < 1e-5, (n, N, K)
        assert int(tick.abs().sum()) == 0


def test_fused_loss_edge_sizes():
    """csrc/loss.cu on batches that are not a multiple of the CTA width, a single ray, and no valid latent row."""
    from robir_b200 import loss as L
    gen = torch.Generator().manual_seed(2)
    for N, n_lat, valid_rows in [(1, 1, 1), (1500, 1500, 700), (1024, 37, 37), (64, 64, 0)]:
        t = lambda *s: torch.rand(*s, generator=gen).cuda().requires_grad_(True)
        sg, ind, alb, albr, r, rr = t(N, 3), t(N, 3), t(N, 3), t(N, 3), t(N, 1), t(N, 1)
        z = (torch.randn(n_lat, 32, generator=gen)).cuda().requires_grad_(True)
        lgt = torch.randn(16, 7, generator=gen).cuda().requires_grad_(True)
        a = torch.tensor(0.01).cuda().requires_grad_(True)
        gt = torch.rand(1, N, 3, generator=gen).cuda()
        mask = (torch.rand(N, generator=gen) > 0.3).cuda()
        zv = (torch.arange(n_lat) < valid_rows).cuda()
        loss, parts = L._FusedPBRLoss.apply(sg, ind, a, alb, albr, r, rr, z, lgt, gt, mask, zv, (1.0, 1.0, 0.1, False))
        # torch restatement (model/loss.py:61-125, train_pbr.py:313-346)
        shift = torch.clamp(torch.clamp(a * 10 + 0.5, 0, 1), 1e-4, 1)
        x = sg + ind
        ldr = x * (2.51 * x + 0.03) / (x * (2.43 * x + 0.59) + 0.14) / shift ** 0.2
        rgb = ((ldr - gt.reshape(-1, 3)).abs() * mask[:, None]).sum() / N
        smooth = (alb - albr).abs().mean() + (r[:, 0] - rr[:, 0]).abs().mean() * 0.2
        rho_hat = (torch.sigmoid(z) * zv[:, None]).sum(0) / zv.sum().clamp(min=1)
        kl = torch.mean(0.05 * torch.log(0.05 / (rho_hat + 1e-4)) + 0.95 * torch.log(0.95 / (1 - rho_hat + 1e-4)))
        ref = rgb + kl + 0.1 * smooth + L.white_loss(lgt)
        assert abs(loss.item() - ref.item()) < 2e-5 * max(1.0, abs(ref.item())), (N, loss.item(), ref.item())
        leaves = [sg, ind, a, alb, albr, r, rr, z, lgt]
        for g1, g2 in zip(torch.autograd.grad(loss, leaves), torch.autograd.grad(ref, leaves)):
            assert torch.isfinite(g1).all()
            assert (g1 - g2).abs().max().item() <= 2e-4 * max(1e-7, g2.abs().max().item()), N


def test_static_step_without_any_hit(synth_sd16):
    """Fixed-capacity path with a camera that sees nothing: zero active rows everywhere, finite loss and gradients."""
    import robir_b200
    from robir_b200 import rng
    from robir_b200.loss import InvLoss, pbr_step_loss
    m = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=16)))
    m.load_state_dict(synth_sd16, strict=True)
    m.cuda().train()
    m.generate()
    m.static_shapes = True
    fn = InvLoss()
    fn.static_shapes = True
    N = 128
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(1, n=N)).items()}
    inp["pose"] = inp["pose"].clone()
    inp["pose"][0, :3, 3] = torch.tensor([5.0, 5.0, 5.0]).cuda()        # looking along -z from far off axis: no hit
    inp["hdr_shift"] = m.gamma.hdr_shift.as_input().expand(N, 1)
    rng.set_mode("device")
    try:
        out = m(inp, trainstage="Material", train_spec=True)
        assert int(out["network_object_mask"].sum()) == 0
        for k in ("sg_rgb", "indir_rgb", "normals", "roughness"):
            assert torch.all(out[k] == 1.0), k
        loss, _ = pbr_step_loss(m, fn, out, {"rgb": torch.full((1, N, 3), 0.5).cuda()})
        loss.backward()
        assert torch.isfinite(loss)
        for p in m.envmap_material_network.parameters():
            assert p.grad is None or torch.isfinite(p.grad).all()
    finally:
        rng.set_mode("cpu")


def test_tc_layer_engine_matches_ffma_chain(model16):
    """csrc/tc_mlp.cu (tcgen05 layer engine, bf16 hi/lo 3-term split) against the exact-fp32 FFMA chain kernel on the
    same networks, points and noise: forward outputs within 1e-4, input / weight gradients within the tensor-core
    gradient bounds; also with an inactive tail (active_rows) and a ragged row count."""
    from robir_b200 import ops, rng
    gen = torch.Generator().manual_seed(17)
    mat, ind = model16.envmap_material_network, model16.indirect_illum_network
    ind.train_weights = True
    try:
        for n, n_act in ((1024, None), (1000, None), (1024, 300)):
            pts = (torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1) * 0.3).cuda()
            hs = torch.full((n, 1), 0.5).cuda().requires_grad_(True)
            res = {}
            for eng in ("ffma", "tc"):
                ops.ENGINE["mlp"] = eng
                for prm in list(mat.parameters()) + list(ind.parameters()):
                    prm.grad = None
                hs.grad = None
                torch.manual_seed(3)
                rng.set_mode("cpu")
                ctx = ops.active_rows(torch.tensor([n_act], dtype=torch.int32).cuda()) if n_act else None
                if ctx:
                    ctx.__enter__()
                try:
                    sgs, env = ind(pts, hs)
                    m = mat(pts, train_spec=True)
                finally:
                    if ctx:
                        ctx.__exit__(None, None, None)
                outs = [sgs, env, m["sg_roughness"], m["sg_diffuse_albedo"], m["sg_normal_map"], m["random_xi_roughness"]]
                live = slice(0, n_act) if n_act else slice(None)
                sum((o[live] * torch.linspace(0.5, 1.5, o[live].numel(), device=o.device).reshape(o[live].shape)).sum()
                    for o in outs if o.requires_grad).backward()
                enc = mat.spec_brdf_encoder_layer.brdf_encoder_layer
                res[eng] = ([o.detach()[live].clone() for o in outs],
                            [hs.grad.clone(), enc[0].weight.grad.clone(), enc[4].weight.grad.clone(),
                             enc[8].bias.grad.clone(), ind.lobe_layer[2].weight.grad.clone()])
            for a, b in zip(res["tc"][0], res["ffma"][0]):
                assert rel_err(a, b) < REL, (n, n_act, rel_err(a, b))
            for a, b in zip(res["tc"][1], res["ffma"][1]):
                grad_close(a, b, 2e-3, 2e-2, engine="tc_bf16")
    finally:
        ops.ENGINE["mlp"] = "tc"
        ind.train_weights = False


def test_static_compact_loss_matches_ray_order_loss(model16):
    """Fixed-capacity forward: the fused loss on the compacted tensors (rows gathered through `order`) against the torch
    restatement on the ray-order outputs of the same forward -- value and gradients."""
    from robir_b200 import loss as L, rng
    model16.generate()
    N = 512
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(12, n=N, crop=460)).items()}
    inp["hdr_shift"] = model16.gamma.hdr_shift.as_input().expand(N, 1)
    gen = torch.Generator().manual_seed(9)
    gt = {"rgb": torch.rand(1, N, 3, generator=gen).cuda()}
    fn = L.InvLoss()
    fn.static_shapes = True
    model16.static_shapes = True
    rng.set_mode("device")
    try:
        out = model16(inp, trainstage="Material", train_spec=True)
        assert 0 < int(out["network_object_mask"].sum()) < N
        got, parts = L.pbr_step_loss(model16, fn, out, gt)
        L.FUSED_LOSS = False
        try:
            ref, ref_parts = L.pbr_step_loss(model16, fn, out, gt)
        finally:
            L.FUSED_LOSS = True
        assert abs(got.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
        mat = model16.envmap_material_network
        enc = mat.spec_brdf_encoder_layer.brdf_encoder_layer
        leaves = [mat.lgtSGs, mat.specular_reflectance, model16.gamma.hdr_shift.adapt_illum, enc[0].bias, enc[8].weight]
        g_ref = torch.autograd.grad(ref, leaves, retain_graph=True)
        g_got = torch.autograd.grad(got, leaves, retain_graph=True)
        for a, b in zip(g_got, g_ref):
            assert (a - b).abs().max().item() <= 5e-4 * max(1e-7, b.abs().max().item())
    finally:
        rng.set_mode("cpu")
        model16.static_shapes = False


# ----------------------------------------------------------------------------------------------------------------------
# CESR extras
# ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture(params=["torch", "tc"])
def wn_engine(request):
    """Engines of the CESR stage's weight-normed 512-wide chains: cuBLAS cross-check and the tcgen05 layer engine."""
    from robir_b200 import ops
    old = ops.ENGINE["wn"]
    ops.ENGINE["wn"] = request.param
    yield request.param
    ops.ENGINE["wn"] = old


def _cesr_nets():
    from robir_b200 import cesr
    sh, nr = synthetic.cesr_state_dicts(0)
    shadow, normal = cesr.WnMLP(191, 2), cesr.WnMLP(63, 3)
    shadow.load_state_dict(sh, strict=True)
    normal.load_state_dict(nr, strict=True)
    return shadow.cuda(), normal.cuda(), sh, nr


@pytest.mark.parametrize("n,N,K", [(5000, 512, 512), (4097, 321, 191), (70, 2, 512), (64, 512, 63), (9000, 3, 512),
                                   (1, 130, 129)])
def test_tl_wgrad_vs_fp64(n, N, K):
    """Tensor-core weight gradients (robir_tl_wgrad: transposed hi/lo images + split-K tcgen05 GEMM + fixed-order
    reduction) against an fp64 matmul: dW = G^T A and db = column sums, ragged rows / columns, strided inputs; two
    launches agree bit for bit."""
    import ctypes
    from robir_b200 import ops
    from robir_b200._lib import lib, check
    gen = torch.Generator().manual_seed(n + N + K)
    Gf = torch.randn(n, N + 5, generator=gen).cuda() * torch.logspace(-6, 0, N + 5).cuda()      # wide dynamic range
    Af = torch.randn(n, K + 3, generator=gen).cuda()
    def run(G, A, rows, n_active=None):
        dW, db = torch.full((N, K), float("nan"), device="cuda"), torch.full((N,), float("nan"), device="cuda")
        work = torch.empty(lib().robir_tl_wgrad_workspace(rows, N, K, ops.sm_count()), dtype=torch.uint8, device="cuda")
        check(lib().robir_tl_wgrad(G.data_ptr(), G.shape[1], A.data_ptr(), A.shape[1], rows, N, K,
                                   n_active.data_ptr() if n_active is not None else None, work.data_ptr(),
                                   dW.data_ptr(), db.data_ptr(), ops.sm_count(), ops.stream()))
        return dW, db
    dW, db = run(Gf, Af, n)
    dW2, db2 = run(Gf, Af, n)
    assert torch.equal(dW, dW2) and torch.equal(db, db2)
    ref = Gf[:, :N].double().T @ Af[:, :K].double()
    refb = Gf[:, :N].double().sum(0)
    # per-row tolerance: every row of dW has its own scale (the logspace above)
    err = ((dW.double() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-300)).max().item()
    errb = ((db.double() - refb).abs() / Gf[:, :N].double().abs().sum(0)).max().item()
    assert err < 2e-5, err
    assert errb < 1e-6, errb
    # fixed-capacity form: the same rows at the front of a larger batch whose tail holds garbage that must not be read
    cap = n + 777
    Gc = torch.full((cap, Gf.shape[1]), float("nan"), device="cuda")
    Ac = torch.full((cap, Af.shape[1]), float("nan"), device="cuda")
    Gc[:n], Ac[:n] = Gf, Af
    dW3, db3 = run(Gc, Ac, cap, torch.tensor([n], dtype=torch.int32, device="cuda"))
    err3 = ((dW3.double() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-300)).max().item()
    assert err3 < 2e-5 and torch.isfinite(db3).all(), err3
    assert ((db3.double() - refb).abs() / Gf[:, :N].double().abs().sum(0)).max().item() < 1e-6


def test_vis_network_training_on_large_batches(model16):
    """VisNetwork logits + all weight gradients on a Vis-stage-sized batch (trace_radiance trains it on n_hit x 512 rows,
    implicit_differentiable_renderer.py:632-634): layer engine with the persistent large-batch kernel and tensor-core weight
    gradients (robir_tl_wgrad) against the exact-fp32 FFMA chain."""
    from robir_b200 import ops
    net = model16.visibility_network
    gen = torch.Generator().manual_seed(44)
    n = 6000
    pts = (torch.randn(n, 3, generator=gen) * 0.4).cuda()
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1).cuda()
    gup = torch.randn(n, 2, generator=gen).cuda()
    old = ops.ENGINE["mlp"]
    res = {}
    try:
        for eng in ("tc", "ffma"):
            ops.ENGINE["mlp"] = eng
            net.zero_grad()
            out = net(pts, dirs)
            (out * gup).sum().backward()
            res[eng] = (out.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()})
    finally:
        ops.ENGINE["mlp"] = old
        net.zero_grad()
    assert rel_err(res["tc"][0], res["ffma"][0]) < 1e-5
    for k, g in res["ffma"][1].items():
        grad_close(res["tc"][1][k], g, 1e-3, 1e-2, engine="tc_bf16")


def test_color_chain_eval_on_layer_engine(model16):
    """borrow_color's colour network (weight-normed 289 -> 256 x4 -> 3, ReLU; model/neus_model.py:440-520 RenderingNetwork) on
    the tensor-core layer engine (evaluation-only path, >= 4096 rows outside autograd) against the FFMA chain kernel and a
    plain torch evaluation of the folded weights."""
    from robir_b200 import ops
    net = model16.implicit_network
    gen = torch.Generator().manual_seed(12)
    pts = (torch.randn(6000, 3, generator=gen) * 0.3).cuda()
    dirs = torch.nn.functional.normalize(torch.randn(6000, 3, generator=gen), dim=-1).cuda()
    old = ops.ENGINE["mlp"]
    try:
        outs = {}
        for eng in ("tc", "ffma"):
            ops.ENGINE["mlp"] = eng
            with torch.no_grad():
                outs[eng] = net.neus_forward(pts, dirs)[0]
        chain = net.__dict__["_color_chain"]
        ops.ENGINE["mlp"] = "tc"
        assert chain.eval_tc_ok(6000, torch.empty(6000, 289)) is False          # autograd on: not the evaluation path
        with torch.no_grad():
            assert chain.eval_tc_ok(6000, torch.empty(6000, 289))
            x = torch.randn(5000, 289, generator=gen).cuda()
            y = chain.eval_tc(x)
            h = x
            cn = net.neus_model.color_network
            for l in range(5):
                lin = getattr(cn, "lin%d" % l)
                W = lin.weight_g * lin.weight_v / lin.weight_v.norm(dim=1, keepdim=True)
                h = torch.nn.functional.linear(h.double(), W.double(), lin.bias.double())
                if l < 4:
                    h = torch.relu(h)
        assert rel_err(y, h.float()) < 2e-5, rel_err(y, h.float())
    finally:
        ops.ENGINE["mlp"] = old
    assert outs["tc"].shape == outs["ffma"].shape == (6000, 3)
    assert (outs["tc"] - outs["ffma"]).abs().max().item() < 2e-5


@pytest.mark.parametrize("n_active", [None, 1536, 0])
def test_tl_layer_big_equals_tile_kernel(n_active):
    """The persistent 128 x 256-tile layer kernel (robir_tl_layer_big: two TMEM accumulators, N = 256 MMAs) against the
    one-CTA-per-tile kernel on the shadow_net chain (K = 191 / 512 / 512 + skip, N = 512 / 321 / 2): outputs and all
    parameter gradients bit for bit, with and without a device-side active-row count (fixed-capacity batches)."""
    from robir_b200 import cesr, ops
    shadow, _, sh, _ = _cesr_nets()
    rows = 4500
    gen = torch.Generator().manual_seed(31)
    x = (torch.randn(rows, 191, generator=gen) * 0.4).cuda()
    gup = torch.randn(rows, 2, generator=gen).cuda()
    na = None if n_active is None else torch.tensor([n_active], dtype=torch.int32, device="cuda")
    if n_active is not None:
        gup[n_active:] = 0                      # contract of the fixed-capacity form
    res = []
    old = ops.TL_BIG_MIN_ROWS, ops.ENGINE["wn"]
    ops.ENGINE["wn"] = "tc"
    try:
        for thr in (1 << 30, 128):
            ops.TL_BIG_MIN_ROWS = thr
            shadow.zero_grad()
            out = cesr.wn_mlp(shadow, x, n_active=na)
            (out * gup).sum().backward()
            res.append((out.detach().clone(), [p.grad.clone() for p in shadow.parameters()]))
    finally:
        ops.TL_BIG_MIN_ROWS, ops.ENGINE["wn"] = old
    (o1, g1), (o2, g2) = res
    assert torch.isfinite(o2).all()
    assert torch.equal(o1, o2)
    if n_active is not None:
        first_inactive = ((n_active + 127) // 128) * 128          # whole row tiles beyond the count are zero-filled
        assert float(o2[first_inactive:(rows // 128) * 128].abs().max()) == 0.0
    for a, b in zip(g1, g2):
        assert torch.isfinite(b).all() and torch.equal(a, b)


@pytest.mark.parametrize("n,N,K", [(5000, 512, 512), (4097, 321, 191), (300, 2, 512), (128, 512, 63), (9000, 3, 512),
                                   (1, 130, 129)])
def test_tl_wgrad_mn_vs_fp64(n, N, K):
    """Weight gradients straight from the layer engine's images (robir_tl_wgrad_mn: the K-major SWIZZLE_128B image blocks
    read as MN-major tcgen05 operands, contraction over the rows, db through a tile of ones) against fp64, on images
    written by robir_tl_pack_rows; ragged shapes, bitwise repeatability, the fixed-capacity form."""
    from robir_b200 import ops
    from robir_b200._lib import lib, check
    gen = torch.Generator().manual_seed(n + N + K + 1)
    Gf = (torch.randn(n, N, generator=gen) * torch.logspace(-6, 0, N)).cuda()
    Af = torch.randn(n, K, generator=gen).cuda()

    def run(G, A, n_active=None):
        rows = G.shape[0]
        gi, ai = ops._tl_rows_image(G, N), ops._tl_rows_image(A, K)
        dW, db = torch.full((N, K), float("nan"), device="cuda"), torch.full((N,), float("nan"), device="cuda")
        work = torch.empty(lib().robir_tl_wgrad_mn_workspace(rows, N, K, ops.sm_count()), dtype=torch.uint8, device="cuda")
        check(lib().robir_tl_wgrad_mn(gi.data_ptr(), (N + 63) // 64, ai.data_ptr(), (K + 63) // 64, rows, N, K,
                                      n_active.data_ptr() if n_active is not None else None, work.data_ptr(),
                                      dW.data_ptr(), db.data_ptr(), ops.sm_count(), ops.stream()))
        return dW, db
    dW, db = run(Gf, Af)
    dW2, db2 = run(Gf, Af)
    assert torch.equal(dW, dW2) and torch.equal(db, db2)
    ref = Gf.double().T @ Af.double()
    refb = Gf.double().sum(0)
    err = ((dW.double() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-300)).max().item()
    errb = ((db.double() - refb).abs() / Gf.double().abs().sum(0).clamp_min(1e-300)).max().item()
    assert err < 2e-5, err
    assert errb < 2e-5, errb                   # db passes through the bf16 hi/lo split as well
    cap = n + 777                              # fixed-capacity form: zero G rows beyond the count, anything in A
    Gc = torch.zeros(cap, N, device="cuda")
    Ac = torch.randn(cap, K, generator=gen).cuda()
    Gc[:n], Ac[:n] = Gf, Af
    dW3, db3 = run(Gc, Ac, torch.tensor([n], dtype=torch.int32, device="cuda"))
    err3 = ((dW3.double() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-300)).max().item()
    assert err3 < 2e-5 and torch.isfinite(db3).all(), err3


@pytest.mark.parametrize("which,rows", [("shadow", 1024), ("shadow", 1000), ("normal", 333), ("shadow", 4500)])
def test_wn_chain_vs_oracle(which, rows, wn_engine):
    """shadow_net / normal_net (weight-normed, softplus(100), skip concat at layer 4) forward + every parameter
    gradient against the oracle's wn_mlp (neus_model.py:385-417) on CPU, ragged row counts included."""
    from robir_b200 import cesr
    shadow, normal, sh, nr = _cesr_nets()
    net, sd = (shadow, sh) if which == "shadow" else (normal, nr)
    gen = torch.Generator().manual_seed(23)
    pts = torch.nn.functional.normalize(torch.randn(rows, 3, generator=gen), dim=-1) * 0.33
    emb = O.pe(pts, 10)
    if which == "shadow":
        lab = torch.eye(128)[torch.randint(0, 128, (rows,), generator=gen)]
        emb = torch.cat([emb, lab], -1)
    gup = torch.randn(rows, 2 if which == "shadow" else 3, generator=gen)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.wn_mlp(sdr, "", emb, prefix_dot=False)
    (ref * gup).sum().backward()
    net.zero_grad()
    out = cesr.wn_mlp(net, emb.cuda())
    (out * gup.cuda()).sum().backward()
    assert out.shape == ref.shape
    assert rel_err(out, ref) < REL, rel_err(out, ref)
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        grad_close(p.grad, sdr[k].grad, 1e-3, 1e-2, engine="tc_bf16" if wn_engine == "tc" else "ffma")


@pytest.fixture(scope="module")
def model128():
    import robir_b200
    m = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=128)))
    m.load_state_dict(synthetic.synthetic_state_dict(0, num_lgt_sgs=128), strict=True)
    m.cuda().train()
    m.generate()
    return m


@pytest.mark.parametrize("case", ["cesr_step_300", "cesr_step", "cesr_step_1200"])
def test_cesr_step_vs_golden(golden, model128, wn_engine, case):
    """IDRNetwork.forward('Material') with the CESR hook bound (train_cesr.py:465-544,588) + the stage's step loss
    (:387-430) + backward against the reference's golden outputs and gradients: warm-up phase at iteration 300 (the MLP
    visibility renders, shadow_net is only supervised, the loss is the supervise term), explore phase at 600 (renders
    with the material network's normal map) and project phase at iteration 1200 (renders with normal_net's normals, so
    the render loss reaches normal_net through d render / d normal: 5-18 % of its gradient in this fixture)."""
    from robir_b200 import cesr, rng
    from robir_b200.loss import InvLoss
    from test_golden import CESR_CASES
    g = golden(case)
    cur_iter, white, explore_iter, proj_iter, smooth_w, kl_w = CESR_CASES[case]
    shadow, normal, _, _ = _cesr_nets()
    hook = cesr.ClusteredAlbedoHook(model128, shadow, normal, white_light=white, explore_iter=explore_iter,
                                    proj_iter=proj_iter, explore_smooth=smooth_w, explore_kl=kl_w, proj_smooth=smooth_w,
                                    proj_kl=kl_w, cur_iter=cur_iter)
    assert hook.prefit_option() == {"cesr_step_300": "warmup", "cesr_step": "explore", "cesr_step_1200": "project"}[case]
    old_hook, old_static = model128.get_sg_render, model128.static_shapes
    model128.get_sg_render, model128.static_shapes = hook.get_sg_render, False
    try:
        model128.zero_grad()
        N = g["pix"].shape[0]
        inp = {k: v.cuda() for k, v in synthetic.camera_inputs(g["pix"]).items()}
        inp["hdr_shift"] = model128.gamma.hdr_shift.as_input().expand(N, 1)
        with rng.replay([g["rnd_%d" % i] for i in range(9)]):
            out = model128(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss, _ = hook.pbr_step(InvLoss(), out, {"rgb": g["gt"]})
    finally:
        model128.get_sg_render, model128.static_shapes = old_hook, old_static
    mask = out["network_object_mask"].cpu()
    assert (mask != g["out_network_object_mask"]).sum() == 0, "tracer mask differs from the reference"
    for k in [k[4:] for k in g if k.startswith("out_") and k != "out_network_object_mask"]:
        assert out[k].shape == g["out_" + k].shape, k
        assert rel_err(out[k], g["out_" + k]) < REL, (k, rel_err(out[k], g["out_" + k]))
    assert abs(loss.item() - g["loss"].item()) < 1e-4 * max(1.0, abs(g["loss"].item()))
    loss.backward()
    mat = model128.envmap_material_network
    eng = "tc_bf16" if wn_engine == "tc" else "ffma"
    checks = []
    if cur_iter > 500:
        checks = [(mat.lgtSGs.grad, g["g_lgtSGs"]), (mat.specular_reflectance.grad, g["g_spec"]),
                  (model128.gamma.hdr_shift.adapt_illum.grad, g["g_adapt"])]
    else:
        assert mat.lgtSGs.grad is None or float(mat.lgtSGs.grad.abs().max()) == 0.0
    checks += [(shadow.lin8.weight_v.grad, g["g_shadow_lin8_v"]), (shadow.lin8.bias.grad, g["g_shadow_lin8_bias"]),
              (shadow.lin4.weight_g.grad, g["g_shadow_lin4_g"]), (shadow.lin0.bias.grad, g["g_shadow_lin0_bias"]),
              (shadow.lin0.weight_v.grad.sum(0), g["g_shadow_lin0_v_colsum"]),
              (normal.lin8.weight_v.grad, g["g_normal_lin8_v"]), (normal.lin0.bias.grad, g["g_normal_lin0_bias"]),
              (normal.lin3.weight_g.grad, g["g_normal_lin3_g"])]
    for a, b in checks:
        grad_close(a, b, 5e-3, 3e-2, eng)


def test_cesr_warmup_phase(model128):
    """Warm-up phase (iteration <= 500): the MLP visibility renders, shadow_net is only supervised, the step loss is the
    supervise term alone."""
    from robir_b200 import cesr
    from robir_b200.loss import InvLoss
    shadow, normal, _, _ = _cesr_nets()
    hook = cesr.ClusteredAlbedoHook(model128, shadow, normal, cur_iter=300)
    assert hook.prefit_option() == "warmup"
    N = 256
    inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(4, n=N, crop=400)).items()}
    inp["hdr_shift"] = model128.gamma.hdr_shift.as_input().expand(N, 1)
    old_hook, old_static = model128.get_sg_render, model128.static_shapes
    model128.get_sg_render, model128.static_shapes = hook.get_sg_render, False
    try:
        out = model128(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        assert int(out["network_object_mask"].sum()) > 0
        loss, _ = hook.pbr_step(InvLoss(), out, {"rgb": torch.full((1, N, 3), 0.4)})
        assert loss is out["gradient_error"] or torch.equal(loss, out["gradient_error"])
        shadow.zero_grad()
        loss.backward()
        assert all(torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0 for p in shadow.parameters())
    finally:
        model128.get_sg_render, model128.static_shapes = old_hook, old_static


def test_neus_stage1_render_vs_golden(model16, synth_sd16):
    """the stage-1 NeuS renderer (render_neus: 64 + 4 x 16 hierarchical depths, render_core
    compositing) on the CUDA kernels vs the unmodified reference's golden outputs (float64 run of
    neus/volume_render/sdf_render.py, tests/golden/make_golden.py) -- evaluation mode and the training-time sampler's
    jittered depths.  Composited quantities agree to the fp32 floor; the individual depths / weights pass through the
    inverse-CDF sampler, which amplifies evaluation-order noise (the fp32 oracle itself sits at 3.5e-4 there)."""
    import neus_stage1 as N1
    from robir_b200 import neus_stage1 as R1
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "neus_stage1.npz"))
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    net = model16.implicit_network
    c = lambda k: g[k].float().cuda()
    with torch.no_grad():
        ev = R1.render_neus(net, c("rays_o"), c("rays_d"), c("near"), c("far"), None, 0.3)
        tr = R1.render_neus(net, c("rays_o"), c("rays_d"), c("near"), c("far"), c("t_rand"), 0.3)
    for k in ("rgb", "dist", "acc"):
        assert rel_err(ev[k], g["eval_" + k].float()) < REL, ("eval", k, rel_err(ev[k], g["eval_" + k].float()))
        assert rel_err(tr[k], g["out_" + k].float()) < REL, ("jittered", k)
    assert abs(float(tr["sim_or_grad"]) - float(g["out_sim_or_grad"])) < REL
    assert tr["weights"].shape == (32, 128) and tr["means"].shape == (32, 128)
    assert (tr["weights"].cpu() - g["out_weights"].float()).abs().max() < 2e-3
    assert (tr["means"].cpu() - g["out_means"].float()).abs().max() < 2e-3
    # the sampler's invariants at a larger size: depths sorted, weights a sub-probability, accumulation = their sum
    gen = torch.Generator().manual_seed(3)
    B = 777
    o = torch.nn.functional.normalize(torch.randn(B, 3, generator=gen), dim=-1) * 3.0
    d = torch.nn.functional.normalize(-o + 0.4 * torch.randn(B, 3, generator=gen), dim=-1)
    near, far = torch.full((B, 1), 1.0), torch.full((B, 1), 5.0)
    with torch.no_grad():
        big = R1.render_neus(net, o.cuda(), d.cuda(), near.cuda(), far.cuda(), torch.rand(B, 1, generator=gen).cuda(), 1.0)
        ref = N1.render_neus({k: v for k, v in synth_sd16.items() if k.startswith("implicit_network.")}, o[:64], d[:64],
                             near[:64], far[:64], None, 1.0, training=False)
        sub = R1.render_neus(net, o[:64].cuda(), d[:64].cuda(), near[:64].cuda(), far[:64].cuda(), None, 1.0)
    assert (big["means"][:, 1:] >= big["means"][:, :-1]).all()
    assert (big["weights"] >= 0).all() and (big["acc"] <= 1.0 + 1e-4).all()
    assert rel_err(big["weights"].sum(-1), big["acc"]) < 1e-5
    for k in ("rgb", "dist", "acc"):
        assert rel_err(sub[k], ref[k]) < REL, ("oracle", k, rel_err(sub[k], ref[k]))
    with pytest.raises(Exception):
        R1.render_neus(net, o[:4].cuda(), d[:4].cuda(), near[:4].cuda(), far[:4].cuda())     # autograd on: refused


@pytest.mark.parametrize("K,max_iter", [(1024, -1), (700, -1), (33, -1), (1500, 32)])
def test_octree_cluster_walk_matches_cooperative_walk(model16, K, max_iter):
    """The single-cluster variant of the walk (ROBIR_OCTREE_CLUSTER=1: hardware cluster barrier + distributed-shared-
    memory live count per lock-step iteration; slower than the cooperative grid on this part, see csrc/trace.cu) against
    the cooperative grid: same per-ray arithmetic, bit-identical distances / hit points / masks / iteration counts."""
    from robir_b200 import ops
    model16.generate()
    tree = model16.ray_tracer.sdf_octree
    gen = torch.Generator().manual_seed(K)
    o = (torch.nn.functional.normalize(torch.randn(K, 3, generator=gen), dim=-1) * (0.9 if max_iter > 0 else 2.0)).cuda()
    d = torch.nn.functional.normalize(-o.cpu() + 0.3 * torch.randn(K, 3, generator=gen), dim=-1).cuda()
    res = {}
    for mode in ("cluster", "cooperative"):
        if mode == "cluster":
            os.environ["ROBIR_OCTREE_CLUSTER"] = "1"
        try:
            x, hit, t, cnt = ops.octree_cast(tree, o, d, max_iter=max_iter, o_div=1, return_stats=True)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("ROBIR_OCTREE_CLUSTER", None)
        res[mode] = (x.cpu(), hit.cpu(), t.cpu(), cnt.cpu())
    a, b = res["cluster"], res["cooperative"]
    assert torch.equal(a[1], b[1]) and 0 < int(a[1].sum()) < K
    assert torch.equal(a[2].nan_to_num(-7.0), b[2].nan_to_num(-7.0)) and torch.equal(a[0].nan_to_num(-7.0), b[0].nan_to_num(-7.0))
    n_it = int(a[3][-6])                                          # counters[kMaxIter + 2] = iterations executed
    assert n_it == int(b[3][-6]) and n_it > 3
    assert torch.equal(a[3][:n_it], b[3][:n_it])                  # live count per iteration


@pytest.mark.parametrize("n", [1, 31, 33, 129, 5000])
def test_sdf_tensor_core_engine_vs_ffma_and_oracle(synth_sd16, model16, n):
    """ENGINE["sdf"] = "tc" (csrc/sdf_tc.cu: the SDF network as a persistent tcgen05 kernel, value + forward-mode normal
    + features) against the FFMA kernel and the oracle: ragged sizes (tile = 32 points with the normal, 128 without),
    both coordinate conventions, the fixed-capacity active-row count."""
    from robir_b200 import ops
    gen = torch.Generator().manual_seed(100 + n)
    pts = (torch.rand(n, 3, generator=gen) * 2 - 1) * 0.7
    w = model16.implicit_network._w
    ref = O.implicit_forward(synth_sd16, pts[:64])
    ref_g = O.implicit_gradient(synth_sd16, pts[:64])[:, 0, :]
    out = {}
    old, old_min = ops.ENGINE["sdf"], ops.SDF_TC_MIN_ROWS
    ops.SDF_TC_MIN_ROWS = 1                       # force the tensor-core kernel at every size
    try:
        for eng in ("ffma", "tc"):
            ops.ENGINE["sdf"] = eng
            a = ops.sdf_eval(w, pts.cuda(), want_grad=True, want_feat=True)
            b = ops.sdf_eval(w, pts.cuda())
            c = ops.sdf_eval(w, pts.cuda(), in_scale=1.0, sdf_scale=1.0, feat_scale=1.0, want_feat=True)
            n_act = torch.tensor([max(n // 2, 0)], dtype=torch.int32, device="cuda")
            with ops.active_rows(n_act):
                d = ops.sdf_eval(w, pts.cuda(), want_grad=True)
            out[eng] = (a, b, c, d)
    finally:
        ops.ENGINE["sdf"], ops.SDF_TC_MIN_ROWS = old, old_min
    (s, g, f), (s2, _, _), (s3, _, f3), (s4, g4, _) = out["tc"]
    (s_f, g_f, f_f), (s2_f, _, _), (s3_f, _, f3_f), (s4_f, g4_f, _) = out["ffma"]
    m = min(n, 64)
    assert rel_err(s[:m], ref[:m, 0]) < REL and rel_err(f[:m], ref[:m, 1:]) < REL and rel_err(g[:m], ref_g[:m]) < REL
    for a, b in ((s, s_f), (g, g_f), (f, f_f), (s2, s2_f), (s3, s3_f), (f3, f3_f)):
        assert a.shape == b.shape and rel_err(a, b) < 2e-5, rel_err(a, b)
    assert torch.equal(s2, s)                       # value-only tiles (128 points) give the same numbers as jet tiles
    k = n // 2
    if k:
        assert rel_err(s4[:k], s_f[:k]) < 2e-5 and rel_err(g4[:k], g_f[:k]) < 2e-5
    assert float(s4[k:].abs().sum()) == 0.0 and float(g4[k:].abs().sum()) == 0.0
