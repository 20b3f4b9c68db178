"""Host-side logic checks: the host+device math headers the CUDA kernels are built from (sg_math.h, octree_walk.h),
compiled here with g++ (tests/hostcheck/hostcheck.cpp -- test infrastructure, never part of the product library) and
compared with the oracle.  Proves the walk/shading/derivative logic on machines without a GPU; the -m gpu tests then
prove the kernels themselves."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import robir_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
SO = os.path.join(HERE, "hostcheck", "libhostcheck.so")
CSRC = os.path.join(os.path.dirname(HERE), "robir_b200", "csrc")


@pytest.fixture(scope="module")
def hc():
    deps = [SRC, os.path.join(CSRC, "sg_math.h"), os.path.join(CSRC, "octree_walk.h"), os.path.join(CSRC, "loss_math.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I" + CSRC, SRC,
                               "-o", SO])
    return ctypes.CDLL(SO)


def fp(t):
    return t.ctypes.data_as(ctypes.c_void_p)


def arr(t, dtype=np.float32):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(dtype))


def test_octree_walk_matches_golden(hc, golden, oracle_octrees):
    from robir_b200.ops import PackedOctree
    g = golden("octree")
    prim, sec = oracle_octrees
    a = prim.arrays()
    tree = PackedOctree(a["root"], a["boxes"], a["non_leaf"], a["links"], a["grid"], a["sdf_val"], a["sdf_grad"],
                        a["min_step"], "cpu")
    nodes, grid, grad = arr(tree.nodes), arr(tree.grid, np.int32), arr(tree.sdf_grad)
    root = np.array(tree.root, dtype=np.float32)

    def cast(o, d, o_div, max_iter):
        o, d = arr(o.reshape(-1, 3)), arr(d.reshape(-1, 3))
        K = d.shape[0]
        t, hit, x = np.empty(K, np.float32), np.empty(K, np.uint8), np.empty((K, 3), np.float32)
        iters = ctypes.c_int()
        hc.hc_octree_cast(fp(nodes), tree.n_nodes, fp(grid), *[int(s) for s in tree.grid.shape], fp(root), fp(grad),
                          fp(o), fp(d), K, o_div, max_iter, ctypes.c_float(tree.refine_limit),
                          ctypes.c_float(tree.last_sdf), fp(t), fp(hit), fp(x), ctypes.byref(iters))
        return torch.from_numpy(t), torch.from_numpy(hit).bool(), torch.from_numpy(x), iters.value

    # (a) bit-exact against the oracle walking the SAME arrays; (b) within fp32 noise of the reference's golden
    # (the reference tree's cached sdf values differ from the oracle's by ~1e-6: weight-norm folding order)
    t, hit, x, iters = cast(g["cam_loc"], g["ray_dirs"], g["ray_dirs"].shape[1], -1)
    xo, ho, to = prim.trace(g["cam_loc"], g["ray_dirs"])
    assert torch.equal(hit, ho) and iters > 10
    assert torch.equal(t[hit], to[hit]), "primary walk must be bit-exact vs. the oracle on the same tree"
    assert torch.equal(x[hit], xo[hit])
    assert torch.equal(hit, g["prim_mask"]) and (t[hit] - g["prim_t"][hit]).abs().max() < 2e-5
    t, hit, x, _ = cast(g["sec_o"], g["sec_d"], g["sec_d"].shape[1], 32)
    xo, ho, to = sec.trace(g["sec_o"], g["sec_d"])
    assert torch.equal(hit, ho) and torch.equal(t, to) and torch.equal(x, xo)
    assert torch.equal(hit, g["sec_mask"]) and (t - g["sec_t"]).abs().max() < 2e-5
    t, hit, x, _ = cast(g["edge_o"], g["edge_d"], 1, -1)
    xo, ho, to = prim.trace(g["edge_o"], g["edge_d"])
    assert hit.tolist() == [False, True] and torch.isnan(t[0]) and t[1] == to[1]
    assert abs(float(t[1]) - float(g["edge_t"][1])) < 2e-5


def _rand_scene(n, M, Mi, seed=0):
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=gen)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    raw = nrm + 0.8 * torch.randn(n, 3, generator=gen)
    raw[::7] = -raw[::7]                                   # some back-facing hits (n.v < 0)
    view = raw / (raw.norm(dim=-1, keepdim=True) + 1e-6)   # the runner's normalisation (train_pbr.py:351): with it
    # ||v|| + 1e-6 == 1 and the half vector of a back-facing point is exactly 0 -> sqrt'(0) must not leak NaN
    from robir_b200.synthetic import synthetic_light_sgs
    lgt = synthetic_light_sgs(gen, M)
    lgt[:, 4:] *= torch.sign(torch.randn(M, 3, generator=gen))      # exercise abs()
    ind = torch.cat([torch.randn(n, Mi, 3, generator=gen), r(n, Mi, 1) * 30 + 0.1, r(n, Mi, 3)], -1)
    return dict(normal=nrm, view=view, rough=r(n, 1) * 0.9 + 0.09, albedo=r(n, 3), spec=torch.tensor([[0.05]]),
                lgt=lgt, ind=ind, lv=r(n, M), bvd=r(n), bvi=r(n), integ=r(n, 3) * 6)


def _oracle_sg(s, leaf_names):
    """Oracle render_with_all_sg with the visibilities injected (VisModel bypassed through monkeypatching)."""
    leaves = {k: s[k].clone().requires_grad_(True) for k in leaf_names}
    v = dict(s)
    v.update(leaves)
    saved = (O.get_diffuse_visibility, O.get_specular_visibility)
    calls = []

    def fake_diffuse(points, normals, vis_fn, lobes, lambdas, ut, up, testing=False, return_aux=False):
        return v["lv"].permute(1, 0), dict(n_query=0)

    def fake_spec(points, normals, viewdirs, vis_fn, lobes, lambdas, ut, up, testing=False, inv=False,
                  return_aux=False):
        calls.append(inv)
        return (v["bvi"] if inv else v["bvd"]), dict(n_query=0)

    O.get_diffuse_visibility, O.get_specular_visibility = fake_diffuse, fake_spec
    try:
        n = s["normal"].shape[0]
        rnd = {k: None for k in ("diff_theta", "diff_phi", "spec_theta", "spec_phi", "ind_theta", "ind_phi")}
        out = O.render_with_all_sg(torch.zeros(n, 3), v["normal"], v["view"], v["lgt"], v["spec"].abs(), v["rough"],
                                   v["albedo"], None, rnd, indir_integral=v["integ"], indir_lgtSGs=v["ind"])
    finally:
        O.get_diffuse_visibility, O.get_specular_visibility = saved
    assert calls == [False, True]
    return out, leaves


def test_sg_render_forward_backward_vs_oracle(hc):
    n, M, Mi = 24, 16, 5
    s = _rand_scene(n, M, Mi)
    names = ["rough", "albedo", "spec", "lgt", "ind", "lv", "bvd", "bvi", "integ", "normal"]
    out, leaves = _oracle_sg(s, names)
    keys = ["sg_rgb", "sg_specular_rgb", "sg_diffuse_rgb", "vis_shadow", "indir_rgb", "indir_specular_rgb",
            "indir_diffuse_rgb"]
    gen = torch.Generator().manual_seed(5)
    gup = [torch.randn(n, 3, generator=gen) for _ in keys]
    gup[3].zero_()
    loss = sum((out[k] * g).sum() for k, g in zip(keys, gup))
    loss.backward()

    o = np.empty((n, 7, 3), np.float32)
    g_out = arr(torch.stack(gup, 1))
    res = {k: np.zeros(tuple(s[k].shape), np.float32) for k in names}
    hc.hc_sg_render(n, M, Mi, fp(arr(s["normal"])), fp(arr(s["view"])), fp(arr(s["rough"])), fp(arr(s["albedo"])),
                    ctypes.c_float(float(s["spec"].abs())), fp(arr(s["lgt"])), fp(arr(s["ind"])), fp(arr(s["lv"])),
                    fp(arr(s["bvd"])), fp(arr(s["bvi"])), fp(arr(s["integ"])), fp(o), fp(g_out), fp(res["lgt"]),
                    fp(res["ind"]), fp(res["lv"]), fp(res["bvd"]), fp(res["bvi"]), fp(res["rough"]),
                    fp(res["albedo"]), fp(res["spec"]), fp(res["integ"]), fp(res["normal"]))
    for j, k in enumerate(keys):
        ref = out[k].detach()
        err = (torch.from_numpy(o[:, j]) - ref).abs().max().item()
        assert err <= 2e-5 * max(1.0, ref.abs().max().item()), (k, err)
    for k in names:
        ref = leaves[k].grad
        got = torch.from_numpy(res[k])
        scale = max(1e-6, ref.abs().max().item())
        assert (got - ref).abs().max().item() <= 2e-4 * scale, (k, (got - ref).abs().max().item(), scale)


def test_sample_dirs_forward_backward_vs_oracle(hc):
    gen = torch.Generator().manual_seed(2)
    M, S = 12, 32
    lobes = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen), dim=-1).requires_grad_(True)
    lam = (torch.rand(M, 1, generator=gen) * 40 + 0.6).requires_grad_(True)     # min < 1 -> live sg_range gradient
    ut, up = torch.rand(M, S, generator=gen), torch.rand(M, S, generator=gen)
    light_dirs, sd = O.diffuse_sample_dirs(lobes, lam, ut, up)
    w = torch.exp(lam.unsqueeze(-2) * (torch.sum(sd * light_dirs, dim=-1, keepdim=True) - 1.0))[..., 0]
    gd, gw = torch.randn(M, S, 3, generator=gen), torch.randn(M, S, generator=gen)
    ((sd * gd).sum() + (w * gw).sum()).backward()

    sharp = torch.clamp(lam.detach()[:, 0], min=1e-4)
    sg_range = float(torch.clamp(sharp.min(), max=1.0))
    dirs, ww = np.empty((M * S, 3), np.float32), np.empty(M * S, np.float32)
    g_af, g_aw = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    g_sh, g_lw, g_r = np.zeros(M, np.float32), np.zeros(M, np.float32), ctypes.c_float()
    hc.hc_sample_dirs(M, S, fp(arr(lobes)), fp(arr(lobes)), fp(arr(sharp)), fp(arr(lam[:, 0])),
                      ctypes.c_float(sg_range), fp(arr(ut)), fp(arr(up)), 1, fp(dirs), fp(ww), fp(arr(gd)),
                      fp(arr(gw)), fp(g_af), fp(g_aw), fp(g_sh), fp(g_lw), ctypes.byref(g_r))
    assert np.abs(dirs.reshape(M, S, 3) - sd.detach().numpy()).max() < 2e-6
    assert np.abs(ww.reshape(M, S) - w.detach().numpy()).max() < 1e-5  # exp(lam*(dot-1)), lam <= 40
    assert np.abs(g_af - lobes.grad.numpy()).max() <= 2e-4 * np.abs(lobes.grad.numpy()).max()
    # lambda gradient = via weight (lam_w) + via sharp (clamp passes) + via sg_range (flows to the arg-min lobe)
    g_lam = torch.from_numpy(g_sh + g_lw)
    g_lam[int(torch.argmin(sharp))] += g_r.value if sharp.min() < 1.0 else 0.0
    ref = lam.grad[:, 0]
    assert (g_lam - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


def test_fused_loss_math_matches_torch(hc):
    """loss_math.h (the per-element math of csrc/loss.cu) against torch autograd of the restated PBR loss
    (model/loss.py:61-125, color_correction.py:31-59, train_pbr.py:313-346): value and every gradient, L1 and L2."""
    from robir_b200.loss import white_loss
    gen = torch.Generator().manual_seed(4)
    for N, n_lat, valid_rows, l2, a0 in [(257, 257, 200, 0, 0.01), (64, 40, 40, 1, -0.03), (5, 5, 0, 0, 0.2)]:
        t = lambda *s: torch.rand(*s, generator=gen).requires_grad_(True)
        sg, ind, alb, albr, r, rr = t(N, 3), t(N, 3), t(N, 3), t(N, 3), t(N), t(N)
        z = torch.randn(n_lat, 32, generator=gen).requires_grad_(True)
        lgt = torch.randn(16, 7, generator=gen).requires_grad_(True)
        a = torch.tensor(a0).requires_grad_(True)
        gt = torch.rand(N, 3, generator=gen)
        mask = torch.rand(N, generator=gen) > 0.3
        zv = torch.arange(n_lat) < valid_rows
        shift = torch.clamp(torch.clamp(a * 10 + 0.5, 0, 1), 1e-4, 1)
        x = sg + ind
        ldr = x * (2.51 * x + 0.03) / (x * (2.43 * x + 0.59) + 0.14) / shift ** 0.2
        diff = ldr - gt
        rgb = ((diff * diff if l2 else diff.abs()) * mask[:, None]).sum() / N
        smooth = (alb - albr).abs().mean() + (r - rr).abs().mean() * 0.2
        rho_hat = (torch.sigmoid(z) * zv[:, None]).sum(0) / zv.sum().clamp(min=1)
        kl = torch.mean(0.05 * torch.log(0.05 / (rho_hat + 1e-4)) + 0.95 * torch.log(0.95 / (1 - rho_hat + 1e-4)))
        ref = rgb + kl + 0.1 * smooth + white_loss(lgt)
        leaves = [sg, a, alb, albr, r, rr, z, lgt]
        g_ref = torch.autograd.grad(ref, leaves)
        outs = dict(losses=np.empty(5, np.float32), g_pred=np.empty((N, 3), np.float32), g_adapt=np.empty(1, np.float32),
                    g_albedo=np.empty((N, 3), np.float32), g_albedo_r=np.empty((N, 3), np.float32),
                    g_rough=np.empty(N, np.float32), g_rough_r=np.empty(N, np.float32),
                    g_z=np.empty((n_lat, 32), np.float32), g_lgt=np.empty((16, 7), np.float32))
        ins = [arr(v) for v in (sg, ind, gt)] + [arr(mask, np.uint8)]
        ins2 = [arr(v) for v in (alb, albr, r, rr, z)] + [arr(zv, np.uint8), arr(lgt)]
        hc.hc_pbr_loss(N, n_lat, 16, l2, *[fp(v) for v in ins], ctypes.c_float(a0), *[fp(v) for v in ins2],
                       ctypes.c_float(1.0), ctypes.c_float(1.0), ctypes.c_float(0.1), ctypes.c_float(0.05),
                       *[fp(v) for v in outs.values()])
        assert abs(outs["losses"][0] - ref.item()) < 2e-5 * max(1.0, abs(ref.item())), (N, outs["losses"], ref.item())
        got = [outs["g_pred"], outs["g_adapt"], outs["g_albedo"], outs["g_albedo_r"], outs["g_rough"], outs["g_rough_r"],
               outs["g_z"], outs["g_lgt"]]
        for g1, g2 in zip(got, g_ref):
            g2 = g2.detach().numpy().reshape(g1.shape)
            assert np.isfinite(g1).all()
            assert np.abs(g1 - g2).max() <= 2e-4 * max(1e-7, np.abs(g2).max()), N
