"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  (see oracle/robir_oracle.py header for the import rules)

Orchestration rows of SURVEY.md section 8: a1 ``IDRNetwork.forward`` (implicit_differentiable_renderer.py:290-479),
a8 ``PBRTrainRunner.get_sg_render`` (training/train_pbr.py:348-396), a13 ``IDRNetwork.trace_radiance`` (:566-650).
Random draws are passed in as a dict in the order of SURVEY.md A.4:

    rnd = { 'indir_noise' [n,64], 'brdf_noise' [n,32], 'normal_noise' [n,60],
            'diff_theta','diff_phi' [M,32], 'spec_theta','spec_phi' [n,8], 'ind_theta','ind_phi' [n,8] }

(n = hit rays).  ``tape_to_rnd`` converts a recorded torch.rand/randn tape into that dict.
"""
import numpy as np
import torch

import robir_oracle as O

RND_KEYS = ["indir_noise", "brdf_noise", "normal_noise", "diff_theta", "diff_phi", "spec_theta", "spec_phi",
            "ind_theta", "ind_phi"]


def tape_to_rnd(tape):
    """tape: list of (name, tensor) in draw order for ONE forward('Material') with an octree tracer."""
    assert len(tape) == len(RND_KEYS), "unexpected number of random draws: %d" % len(tape)
    return {k: t for k, (_, t) in zip(RND_KEYS, tape)}


def draw_rnd(n_hit, M, gen=None, S=32):
    """Draw the randoms of one forward('Material') in reference order with a CPU generator."""
    r, rn = (lambda *s: torch.rand(*s, generator=gen)), (lambda *s: torch.randn(*s, generator=gen))
    return dict(indir_noise=rn(n_hit, 64), brdf_noise=rn(n_hit, 32), normal_noise=rn(n_hit, 60),
                diff_theta=r(M, S), diff_phi=r(M, S), spec_theta=r(n_hit, 8), spec_phi=r(n_hit, 8),
                ind_theta=r(n_hit, 8), ind_phi=r(n_hit, 8))


def pbr_get_sg_render(sd, points, view_dirs, indir_lgtSGs, indir_integral, rnd, no_normal=True, is_training=True,
                      stats=None):
    """training/train_pbr.py:348-396."""
    view_dirs = view_dirs / (torch.norm(view_dirs, dim=-1, keepdim=True) + 1e-6)
    normals = O.implicit_gradient(sd, points)[:, 0, :]
    normals = normals / torch.clamp(torch.norm(normals, dim=-1, keepdim=True), 1e-4)
    mat = O.envmap_material(sd, points, rnd["brdf_noise"], rnd["normal_noise"])
    indir_integral = indir_integral * 2 * np.pi
    vis_fn = lambda p, d: O.vis_network(sd, p, d)
    sg = O.render_with_all_sg(points.detach(), normals.detach() if no_normal else mat["sg_normal_map"].detach(),
                              view_dirs, mat["sg_lgtSGs"], mat["sg_specular_reflectance"].abs(), mat["sg_roughness"],
                              mat["sg_diffuse_albedo"], vis_fn, rnd, indir_integral=indir_integral,
                              indir_lgtSGs=indir_lgtSGs, lin_diff=False, testing=not is_training, stats=stats)
    ret = {"normals": normals}
    ret.update(sg)
    ret.update(diffuse_albedo=mat["sg_diffuse_albedo"], roughness=mat["sg_roughness"], metallic=mat["sg_metallic"],
               normal_map=mat["sg_normal_map"], random_xi_roughness=mat["random_xi_roughness"],
               random_xi_metallic=mat["random_xi_metallic"],
               random_xi_diffuse_albedo=mat["random_xi_diffuse_albedo"])
    return ret


def white_loss(lgtSGs):
    """training/train_pbr.py:313-317 == training/train_cesr.py:460-463."""
    lgt = torch.abs(lgtSGs[..., -3:])
    mu = lgt.norm(dim=-1, keepdim=True) + 1e-4
    return (lgt / mu).var(-1).mean() * 0.01


def cesr_prefit_option(cur_iter, explore_iter, proj_iter):
    """training/train_cesr.py:546-559 (is_explore_step / prefit_option)."""
    explore = cur_iter > 500 and cur_iter % (explore_iter + proj_iter) >= proj_iter
    if not explore:
        return "warmup" if cur_iter <= 500 else "project"
    return "explore"


def cesr_get_sg_render(sd, sd_shadow, sd_normal, points, view_dirs, indir_lgtSGs, indir_integral, rnd, cur_iter=600,
                       prefit="explore", white_light=True, is_training=True, stats=None):
    """ClusteredAlbedoTrainRunner.get_sg_render (training/train_cesr.py:465-544).  sd_shadow / sd_normal are the state
    dicts of shadow_net = SDFNetwork(63+128, 2, 512, 8, [4], 0) and normal_net = SDFNetwork(63, 3, 512, 8, [4], 0)
    (:106-110), keys ``lin{l}.weight_g|weight_v|bias``.  rnd as for the PBR hook but diff_theta / diff_phi are [M,8]."""
    view_dirs = view_dirs / (torch.norm(view_dirs, dim=-1, keepdim=True) + 1e-6)
    normals = O.implicit_gradient(sd, points)[:, 0, :]
    normals = normals / torch.clamp(torch.norm(normals, dim=-1, keepdim=True), 1e-4)
    mat = O.envmap_material(sd, points, rnd["brdf_noise"], rnd["normal_noise"])
    lgtSGs = mat["sg_lgtSGs"]
    M = lgtSGs.shape[0]
    assert M == 128, "the CESR hook hard-codes 128 light lobes (train_cesr.py:492-493)"
    indir_integral = indir_integral * 2 * np.pi
    normal_map = mat["sg_normal_map"].detach()
    emb = O.pe(points.detach(), 10)                                                    # get_embedder(10), :106
    shadow_in = torch.cat([emb[:, None, :].expand(-1, 128, -1), torch.eye(128)[None].expand(emb.shape[0], -1, -1)], -1)
    with torch.set_grad_enabled(is_training and torch.is_grad_enabled()):
        diffuse_vis = O.wn_mlp(sd_shadow, "", shadow_in.reshape(-1, shadow_in.shape[-1]), prefix_dot=False)
        normal_new = O.wn_mlp(sd_normal, "", emb, prefix_dot=False)
    normal_new = normal_new / torch.clamp(normal_new.norm(dim=-1, keepdim=True), 1e-4)
    diffuse_vis = torch.softmax(diffuse_vis, -1)[..., 1]
    vis_fn = lambda p, d: O.vis_network(sd, p, d)
    sg = O.render_with_all_sg(points.detach(), normal_new if cur_iter > 1000 else normal_map, view_dirs, lgtSGs,
                              mat["sg_specular_reflectance"].abs(), mat["sg_roughness"], mat["sg_diffuse_albedo"],
                              vis_fn, rnd, indir_integral=indir_integral, indir_lgtSGs=indir_lgtSGs, lin_diff=True,
                              testing=not is_training, stats=stats, diffuse_vis=diffuse_vis, prefit=prefit)
    albedo = mat["sg_diffuse_albedo"]
    sg["sg_rgb"] = sg["sg_diffuse_rgb"] * albedo / np.pi + sg["sg_specular_rgb"]
    sg["indir_rgb"] = sg["indir_diffuse_rgb"] * albedo / np.pi + sg["indir_specular_rgb"]
    supervise = sg["supervise"]
    if white_light and prefit != "warmup":
        supervise = supervise + white_loss(lgtSGs)
    supervise = supervise + ((normal_map - normal_new) ** 2).mean()
    ret = {"normals": normals}
    ret.update(sg)
    ret.update(diffuse_albedo=albedo, roughness=mat["sg_roughness"], metallic=mat["sg_metallic"],
               normal_map=normal_new, gradient_error=supervise, random_xi_roughness=mat["random_xi_roughness"],
               random_xi_metallic=mat["random_xi_metallic"],
               random_xi_diffuse_albedo=mat["random_xi_diffuse_albedo"])
    return ret


def idr_forward(sd, inp, tracer, rnd, trainstage="Material", no_normal=True, is_training=True, stats=None, hook=None):
    """IDRNetwork.forward (implicit_differentiable_renderer.py:290-479), camera-input branch, hdr_shift present.
    tracer(cam_loc[B,3], object_mask[B*N], ray_dirs[B,N,3]) -> points, mask, dists   (no_grad).
    rnd may be a dict or a callable n_hit -> dict (the shapes depend on the hit count).
    hook(points, view_dirs, indir_lgtSGs, indir_integral, rnd) replaces the PBR-stage get_sg_render (:400-409), e.g.
    a closure over cesr_get_sg_render."""
    uv, pose, K = inp["uv"], inp["pose"], inp["intrinsics"]
    object_mask = inp["object_mask"].reshape(-1)
    ray_dirs, cam_loc = O.camera_rays(uv, pose, K)
    B, N, _ = ray_dirs.shape
    with torch.no_grad():
        _, net_mask, dists = tracer(cam_loc, object_mask, ray_dirs)
    points = (cam_loc.unsqueeze(1) + dists.reshape(B, N, 1) * ray_dirs).reshape(-1, 3)
    sdf_output = O.implicit_forward(sd, points)[:, 0:1]
    ray_dirs = ray_dirs.reshape(-1, 3)
    out = dict(points=points, sdf_output=sdf_output, network_object_mask=net_mask, object_mask=object_mask,
               ray_dirs=ray_dirs)
    sm = net_mask
    n_hit = int(sm.sum())
    if callable(rnd):
        rnd = rnd(n_hit)
    total = points.shape[0]
    indirect_sgs = torch.ones(total, 24, 7)
    indirect_sgs[:, :, -3:] = 0
    indirect_integral = torch.ones(total, 3)
    if n_hit > 0:
        sgs, integ = O.indirect_illum(sd, points[sm], inp["hdr_shift"][sm], rnd["indir_noise"])
        indirect_sgs = indirect_sgs.clone()
        indirect_sgs[sm] = sgs
        indirect_integral[sm] = integ
    out["hdr_shift"] = inp["hdr_shift"]
    if trainstage == "Illum":
        out.update(indirect_sgs=indirect_sgs, indir_integral=indirect_integral)
        normals = torch.ones_like(points)
        if n_hit > 0:
            m = O.envmap_material(sd, points[sm], None, rnd["normal_noise"], train_norm=True)
            normals[sm] = m["sg_normal_map"]
        out["normals"] = normals
        return out

    ones3 = lambda: torch.ones(total, 3)
    ones1 = lambda: torch.ones(total, 1)
    buf = dict(sg_rgb=ones3(), indir_rgb=ones3(), sg_diffuse_rgb=ones3(), sg_specular_rgb=ones3(),
               indir_diffuse_rgb=ones3(), indir_specular_rgb=ones3(), normals=ones3(), diffuse_albedo=ones3(),
               roughness=ones3(), metallic=ones1(), normal_map=ones3(), vis_shadow=ones3(),
               random_xi_diffuse_albedo=ones3(), random_xi_roughness=ones3(), random_xi_metallic=ones1())
    gradient_error = torch.tensor(0.0)
    if n_hit > 0:
        if hook is not None:
            ret = hook(points[sm], -ray_dirs[sm], indirect_sgs[sm], indirect_integral[sm], rnd)
        else:
            ret = pbr_get_sg_render(sd, points[sm], -ray_dirs[sm], indirect_sgs[sm], indirect_integral[sm], rnd,
                                    no_normal=no_normal, is_training=is_training, stats=stats)
        if "gradient_error" in ret:
            gradient_error = gradient_error + ret["gradient_error"]                    # :446-447
        for k in buf:
            v = ret[k]
            if k in ("roughness", "random_xi_roughness"):
                v = v.expand(-1, 3)
            b = buf[k].clone()
            b[sm] = v
            buf[k] = b
    out.update(final_t=ones1(), gradient_error=gradient_error, acc=ones1(), bg_rgb=ones3(), surface_mask=sm)
    out.update(buf)
    return out


def trace_radiance(sd, inp, sec_tracer, u, t, nsamp):
    """IDRNetwork.trace_radiance (implicit_differentiable_renderer.py:566-650).  u, t: the two uniform [n_hit*nsamp]
    draws of spherical_uniform (:583-589).  sec_tracer: the max_iter=32 octree tracer (cam_loc [K,3], dirs [K,S,3])."""
    points, hdr_shift, pm = inp["points"], inp["hdr_shift"], inp["network_object_mask"]
    N = points.shape[0]
    trace_rad = torch.zeros(N, nsamp, 3)
    sec_o = points[pm].clone()
    n = sec_o.shape[0]
    sample_dirs = torch.zeros(n, nsamp, 3)
    gt_vis = torch.zeros(N, nsamp, 1).bool()
    pred_vis = torch.zeros(N, nsamp, 2)
    indir_mask = torch.zeros_like(gt_vis)
    gt_integral = torch.zeros_like(points)
    if n > 0:
        uu = u * 2 - 1
        tt = t * torch.pi * 2
        sample_dirs = torch.stack([(1 - uu ** 2) ** 0.5 * torch.cos(tt), (1 - uu ** 2) ** 0.5 * torch.sin(tt), uu],
                                  -1).view(n, nsamp, 3)
        normals = inp["normals"].detach()[pm][:, None, :]
        normals = normals / torch.clamp(torch.norm(normals, dim=-1, keepdim=True), 1e-4)
        back = (normals * sample_dirs).sum(-1) < 0
        with torch.no_grad():
            sec_pts, sec_mask, _ = sec_tracer(sec_o + normals[:, 0] * 0.005, None, sample_dirs)
        if sec_mask.any():
            hp = sec_pts[sec_mask]
            hv = -sample_dirs.reshape(-1, 3)[sec_mask]
            rad = torch.zeros_like(sec_pts)
            rad[sec_mask] = O.batch_borrow_color(sd, hp, hv)
            shift = hdr_shift[pm][:, None, :].expand(-1, nsamp, 1).reshape(-1, 1)
            rad[sec_mask] = O.ldr2hdr(rad[sec_mask] ** 2.2, shift[sec_mask])
            rad = rad.reshape(n, nsamp, 3)
            rad[back] = 0.0
            trace_rad[pm] = rad
        in_p = sec_o.unsqueeze(1).expand(-1, nsamp, 3)
        pred_vis[pm] = O.vis_network(sd, in_p.reshape(-1, 3), sample_dirs.reshape(-1, 3)).reshape(-1, nsamp, 2)
        gt_vis[pm] = sec_mask.reshape(n, nsamp, 1)
        indir_mask[pm] = ~back[..., None] & gt_vis[pm]
        cos_dot = trace_rad[pm] * torch.relu((normals * sample_dirs).sum(-1, keepdims=True))
        hemi = (~back).sum(-1)[..., None]
        gt_integral[pm] = cos_dot.sum(-2) / torch.clamp(hemi, 1e-4)
    return dict(trace_radiance=trace_rad, sample_dirs=sample_dirs, gt_vis=gt_vis, pred_vis=pred_vis,
                indir_mask=indir_mask[..., 0], gt_integral=gt_integral)
