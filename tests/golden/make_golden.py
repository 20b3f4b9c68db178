"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference through oracle/ref_shim.py) on
seeded synthetic weights (robir_b200.synthetic).  Run in the build container only:

    python tests/golden/make_golden.py

The files hold inputs (incl. every random draw) and the reference's outputs; weights are re-created from the seed by
``synthetic_state_dict`` so the fixtures stay small.  tests/test_golden.py checks the oracle against them on any
machine; the -m gpu tests check the CUDA product against them on the B200.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_shim  # noqa: E402
from robir_b200 import synthetic  # noqa: E402

M = 16
SEED = 0


def npy(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def build(use_octree=True, n_steps=100, perturb=0.0):
    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=M, perturb=perturb)
    model = ref_shim.build_reference_model(synthetic.neus_checkpoint_from(sd), num_lgt_sgs=M, use_octree=use_octree,
                                           n_steps=n_steps)
    model.load_state_dict(sd, strict=True)
    model.train()
    ref_shim.bind_pbr_runner(model)
    return sd, model


CESR_CASES = {   # file -> runner settings (confs_sg/hotdog.conf:34-43 explore schedule; confs_sg/truck.conf project schedule)
    "cesr_step_300": dict(cur_iter=300, white_light=True, explore_iter=1000, proj_iter=0),      # warm-up phase
    "cesr_step": dict(cur_iter=600, white_light=True, explore_iter=1000, proj_iter=0, explore_smooth=0.1, explore_kl=1.0),
    "cesr_step_1200": dict(cur_iter=1200, white_light=False, explore_iter=0, proj_iter=1000, proj_smooth=0.001,
                           proj_kl=0.01),
}


def cesr_golden(g, name="cesr_step"):
    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=128)
    model = ref_shim.build_reference_model(synthetic.neus_checkpoint_from(sd), num_lgt_sgs=128)
    model.load_state_dict(sd, strict=True)
    model.train()
    model.ray_tracer.generate(lambda x: model.implicit_network(x)[:, 0], None)
    from model.loss import InvLoss
    runner = ref_shim.bind_cesr_runner(model, **CESR_CASES[name])
    runner.loss = InvLoss(1.0, 0.1, 100.0, 50.0, 1.0, 1.0, 1.0)
    sh, nr = synthetic.cesr_state_dicts(SEED)
    runner.shadow_net.load_state_dict(sh, strict=True)
    runner.normal_net.load_state_dict(nr, strict=True)
    N = 48
    pix = synthetic.training_pixels(3, n=N, crop=400)
    inp = synthetic.camera_inputs(pix)
    gt = torch.rand(1, N, 3, generator=g)
    torch.manual_seed(1234)
    with ref_shim.ReplayRandom() as rec:
        i2 = dict(inp)
        i2["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(N, 1)
        out = model(i2, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss, _ = runner.pbr_step(out, {"rgb": gt})
    for m in (model, runner.shadow_net, runner.normal_net):
        m.zero_grad()
    loss.backward()
    keep = ["points", "network_object_mask", "sg_rgb", "indir_rgb", "sg_diffuse_rgb", "sg_specular_rgb",
            "indir_diffuse_rgb", "indir_specular_rgb", "normals", "diffuse_albedo", "roughness", "normal_map",
            "vis_shadow", "gradient_error"]
    d = {"out_" + k: out[k] for k in keep}
    d.update({"rnd_%d" % i: t for i, (_, t) in enumerate(rec.tape)})
    mat = model.envmap_material_network
    d.update(pix=pix, gt=gt, loss=loss, g_lgtSGs=mat.lgtSGs.grad, g_spec=mat.specular_reflectance.grad,
             g_adapt=model.gamma.hdr_shift.adapt_illum.grad,
             g_shadow_lin8_v=runner.shadow_net.lin8.weight_v.grad, g_shadow_lin8_bias=runner.shadow_net.lin8.bias.grad,
             g_shadow_lin4_g=runner.shadow_net.lin4.weight_g.grad, g_shadow_lin0_bias=runner.shadow_net.lin0.bias.grad,
             g_shadow_lin0_v_colsum=runner.shadow_net.lin0.weight_v.grad.sum(0),
             g_normal_lin8_v=runner.normal_net.lin8.weight_v.grad, g_normal_lin0_bias=runner.normal_net.lin0.bias.grad,
             g_normal_lin3_g=runner.normal_net.lin3.weight_g.grad)
    d = {k: v for k, v in d.items() if v is not None}     # warm-up: the loss is the supervise term, no material gradients
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **npy(d))
    print(name, runner.prefit_option(), ": hits", int(out["network_object_mask"].sum()), "loss", float(loss), "supervise",
          float(out["gradient_error"]))


def stage1_golden():
    """neus/volume_render/sdf_render.py render_neus (unmodified file, loaded by
    ref_shim.load_stage1_renderer) on the synthetic stage-1 weights, in float64 so that the fixture pins the algorithm
    rather than float32 evaluation order (the inverse-CDF sampler amplifies rounding noise by up to 1e5)."""
    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=M)     # the standard (float32-drawn) weights, widened below
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        R = ref_shim.load_stage1_renderer()
        from model.neus_model import NeuSModel
        model = NeuSModel(mode="idr", hashing=False, embed="PE")
        model.load_state_dict(synthetic.neus_checkpoint_from(sd), strict=True)
        model.double()
        gen = torch.Generator().manual_seed(21)
        B = 32
        rays_o = torch.tensor([[0.026, 0.042, 4.0]]).expand(B, 3) + 0.02 * torch.randn(B, 3, generator=gen)
        target = torch.nn.functional.normalize(torch.randn(B, 3, generator=gen), dim=-1) * 0.7
        rays_d = torch.nn.functional.normalize(target - rays_o, dim=-1)
        near, far = torch.full((B, 1), 2.5), torch.full((B, 1), 5.5)
        rays = R.Rays(rays_o, rays_d, rays_d, None, torch.ones(B, 1), near, far)
        gt = torch.rand(B, 3, generator=gen)
        t_rand = torch.rand(B, 1, generator=gen)
        with ref_shim.ReplayRandom(tape=[("rand", t_rand)]):
            ret = R.render_neus(rays, model, 0.3, n_outside=0, white_bkgd=True, is_eval=False)
        loss = ((ret["rgb"] - gt) ** 2).mean() + 0.1 * ret["sim_or_grad"]
        model.zero_grad()
        loss.backward()
        with torch.no_grad():
            ev = R.render_neus(rays, model, 0.3, n_outside=0, white_bkgd=True, is_eval=True)
        d = {"out_" + k: v for k, v in ret.items()}
        d.update({"eval_" + k: ev[k] for k in ("rgb", "dist", "acc")})
        d.update(rays_o=rays_o, rays_d=rays_d, near=near, far=far, gt=gt, t_rand=t_rand, loss=loss,
                 g_sdf_lin8_v_row0=model.sdf_network.lin8.weight_v.grad[0],
                 g_sdf_lin8_v_rowsum=model.sdf_network.lin8.weight_v.grad.sum(1), g_sdf_lin0_bias=model.sdf_network.lin0.bias.grad,
                 g_sdf_lin4_g=model.sdf_network.lin4.weight_g.grad, g_col_lin4_v=model.color_network.lin4.weight_v.grad,
                 g_variance=model.deviation_network.variance.grad)
        np.savez_compressed(os.path.join(HERE, "neus_stage1.npz"), **npy(d))
        print("stage1: acc max", float(ret["acc"].max()), "loss", float(loss))
    finally:
        torch.set_default_dtype(old)


def main():
    torch.set_num_threads(8)
    sd, model = build()
    sdf_fn = lambda x: model.implicit_network(x)[:, 0]
    model.ray_tracer.generate(sdf_fn, None)
    model.octree_ray_tracer.generate(sdf_fn, None)
    oc = model.ray_tracer.sdf_octree
    fp = dict(n_nodes=oc.octree.boxes.shape[0], links_sum=int(oc.octree.links.sum()),
              non_leaf_sum=int(oc.octree.non_leaf.sum()), hit_sum=int(oc.hit_ptr.sum()),
              sdf_val_sum=float(oc.sdf_val.double().sum()), min_step=oc.min_step)
    print("octree fingerprint", fp)

    # ---------------- 1. small nets on 96 points near the surface
    g = torch.Generator().manual_seed(1)
    dirs = torch.nn.functional.normalize(torch.randn(96, 3, generator=g), dim=-1)
    pts = dirs * (0.33 + 0.01 * torch.randn(96, 1, generator=g))
    vdirs = torch.nn.functional.normalize(torch.randn(96, 3, generator=g), dim=-1)
    hs = torch.rand(96, 1, generator=g)
    noise = dict(indir=torch.randn(96, 64, generator=g), brdf=torch.randn(96, 32, generator=g),
                 nrm=torch.randn(96, 60, generator=g))
    with torch.no_grad():
        f = model.implicit_network(pts)
    grad = model.implicit_network.gradient(pts.clone())[:, 0].detach()
    with ref_shim.ReplayRandom(tape=[("randn", noise["indir"])]):
        sgs, env = model.indirect_illum_network(pts, hs)
    with ref_shim.ReplayRandom(tape=[("randn", noise["brdf"]), ("randn", noise["nrm"])]):
        mat = model.envmap_material_network(pts, train_spec=True)
    vis_logits = model.visibility_network(pts, vdirs)
    with torch.no_grad():
        col = model.implicit_network.batch_borrow_color(pts, vdirs)
    np.savez_compressed(os.path.join(HERE, "nets.npz"), **npy(dict(
        pts=pts, vdirs=vdirs, hdr_shift=hs, noise_indir=noise["indir"], noise_brdf=noise["brdf"],
        noise_nrm=noise["nrm"], sdf_feat_head=f[:, :8], sdf_feat_sum=f.sum(-1), grad=grad, indir_sgs=sgs,
        indir_env=env, roughness=mat["sg_roughness"], albedo=mat["sg_diffuse_albedo"], metallic=mat["sg_metallic"],
        normal_map=mat["sg_normal_map"], xi_roughness=mat["random_xi_roughness"],
        xi_albedo=mat["random_xi_diffuse_albedo"], vis_logits=vis_logits, borrow_color=col)))

    # ---------------- 2. octree casts (primary max_iter=-1, secondary max_iter=32)
    pix = synthetic.training_pixels(0, n=512, crop=420)
    inp = synthetic.camera_inputs(pix)
    from utils import rend_util
    rd, cl = rend_util.get_camera_params(inp["uv"], inp["pose"], inp["intrinsics"])
    with torch.no_grad():
        p1, m1, t1 = model.ray_tracer(sdf=sdf_fn, cam_loc=cl, object_mask=inp["object_mask"].reshape(-1),
                                      ray_directions=rd)
        so = pts[:64] + 0.005 * torch.nn.functional.normalize(pts[:64], dim=-1)
        sdir = torch.nn.functional.normalize(torch.randn(64, 8, 3, generator=g), dim=-1)
        p2, m2, t2 = model.octree_ray_tracer(sdf=sdf_fn, cam_loc=so, object_mask=None, ray_directions=sdir)
        # NaN edge case: axis-aligned ray starting on a grid plane
        eo = torch.tensor([[0.0, 0.0, 2.0], [0.05, 0.1, 2.0]])
        ed = torch.tensor([[[0.0, 0.0, -1.0]], [[0.0, 0.0, -1.0]]])
        p3, m3, t3 = model.ray_tracer(sdf=sdf_fn, cam_loc=eo, object_mask=None, ray_directions=ed)
    np.savez_compressed(os.path.join(HERE, "octree.npz"), **npy(dict(
        pix=pix, ray_dirs=rd, cam_loc=cl, prim_points=p1, prim_mask=m1, prim_t=t1, sec_o=so, sec_d=sdir,
        sec_points=p2, sec_mask=m2, sec_t=t2, edge_o=eo, edge_d=ed, edge_mask=m3, edge_t=t3,
        **{"fp_" + k: v for k, v in fp.items()})))
    print("octree: prim hits", int(m1.sum()), "/ 512; sec hits", int(m2.sum()), "/ 512; edge", m3.tolist(), t3.tolist())

    # ---------------- 3. PBR forward + loss + backward (N=160 rays), octree tracer
    pix = synthetic.training_pixels(1, n=160, crop=400)
    inp = synthetic.camera_inputs(pix)
    gt = torch.rand(1, 160, 3, generator=g)
    from model.loss import InvLoss
    model.get_sg_render.__self__.loss = InvLoss(1.0, 0.1, 100.0, 50.0, 1.0, 1.0, 1.0)
    runner = model.get_sg_render.__self__
    torch.manual_seed(1234)
    with ref_shim.ReplayRandom() as rec:
        i2 = dict(inp)
        i2["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(160, 1)
        out = model(i2, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss, _ = runner.pbr_step(out, {"rgb": gt})
    model.zero_grad()
    loss.backward()
    mat = model.envmap_material_network
    dec = mat.spec_brdf_encoder_layer.brdf_decoder_layer
    enc = mat.spec_brdf_encoder_layer.brdf_encoder_layer
    keep = ["points", "network_object_mask", "sg_rgb", "indir_rgb", "sg_diffuse_rgb", "sg_specular_rgb",
            "indir_diffuse_rgb", "indir_specular_rgb", "normals", "diffuse_albedo", "roughness", "metallic",
            "normal_map", "vis_shadow", "random_xi_roughness", "random_xi_diffuse_albedo", "sdf_output"]
    d = {"out_" + k: out[k] for k in keep}
    d.update({"rnd_%d" % i: t for i, (_, t) in enumerate(rec.tape)})
    d.update(pix=pix, gt=gt, loss=loss, g_lgtSGs=mat.lgtSGs.grad, g_spec=mat.specular_reflectance.grad,
             g_adapt=model.gamma.hdr_shift.adapt_illum.grad, g_dec4_bias=dec[4].bias.grad,
             g_dec4_weight=dec[4].weight.grad, g_enc0_bias=enc[0].bias.grad, g_enc8_weight_sum=enc[8].weight.grad.sum(0))
    np.savez_compressed(os.path.join(HERE, "pbr_step.npz"), **npy(d))
    print("pbr: hits", int(out["network_object_mask"].sum()), "loss", float(loss))

    # ---------------- 3b. CESR step (M = 128 lobes, N = 48 rays, explore phase, iteration 600): the hook of
    # training/train_cesr.py:465-544 with seeded shadow_net / normal_net weights, loss of :387-430, backward
    for name in CESR_CASES:
        cesr_golden(torch.Generator().manual_seed(11), name)
    stage1_golden()

    # ---------------- 4. Illum forward + trace_radiance (N=48, nsamp=16)
    model.zero_grad()
    pix = synthetic.training_pixels(2, n=48, crop=360)
    inp = synthetic.camera_inputs(pix)
    torch.manual_seed(99)
    with ref_shim.ReplayRandom() as rec2:
        i2 = dict(inp)
        i2["hdr_shift"] = torch.rand(48, 1)
        o_ill = model(i2, trainstage="Illum")
        tr = model.trace_radiance(o_ill, nsamp=16)
    d = {"rnd_%d" % i: t for i, (_, t) in enumerate(rec2.tape)}
    d.update(pix=pix, indirect_sgs=o_ill["indirect_sgs"], indir_integral=o_ill["indir_integral"],
             normals=o_ill["normals"], points=o_ill["points"], mask=o_ill["network_object_mask"])
    d.update({"tr_" + k: v for k, v in tr.items()})
    np.savez_compressed(os.path.join(HERE, "vis_stage.npz"), **npy(d))
    print("vis stage: hits", int(o_ill["network_object_mask"].sum()), "sec hits", int(tr["gt_vis"].sum()))

    # ---------------- 5. IDR sphere tracer (use_octree=False), eval and training
    sd2, model2 = build(use_octree=False, n_steps=32)
    sdf2 = lambda x: model2.implicit_network(x)[:, 0]
    pix = synthetic.training_pixels(3, n=256, crop=420)
    inp = synthetic.camera_inputs(pix)
    rd, cl = rend_util.get_camera_params(inp["uv"], inp["pose"], inp["intrinsics"])
    om = torch.rand(256, generator=g) > 0.3
    d = dict(pix=pix, object_mask=om)
    for training in (False, True):
        model2.ray_tracer.train(training)
        torch.manual_seed(5)
        uni = torch.empty(32).uniform_(0.0, 1.0)
        torch.manual_seed(5)
        with torch.no_grad():
            p, m, t = model2.ray_tracer(sdf=sdf2, cam_loc=cl, object_mask=om, ray_directions=rd)
        tag = "train" if training else "eval"
        d.update({tag + "_points": p, tag + "_mask": m, tag + "_t": t, "uniform": uni})
        print("raytracing", tag, "hits", int(m.sum()))
    np.savez_compressed(os.path.join(HERE, "raytracing.npz"), **npy(d))


if __name__ == "__main__":
    if sys.argv[1:] == ["stage1"]:
        stage1_golden()
    elif sys.argv[1:] == ["cesr"]:
        torch.set_num_threads(8)
        for name in CESR_CASES:
            cesr_golden(torch.Generator().manual_seed(11), name)
    else:
        main()
