"""Drop-in replacements for model/sg_render.py's public functions with the reference signatures
(get_diffuse_visibility :111, get_specular_visibility :198, render_with_all_sg :304, render_with_sg :343), backed by
the fused CUDA operators.  ``VisModel`` must expose the reference VisNetwork layout (a ``vis_layer`` nn.Sequential of
five Linear layers, 126->256x4->2); an arbitrary callable (e.g. the reference's OctreeVisModel) is rejected loudly.
"""
import numpy as np
import torch

from . import dist as rdist
from . import ops, rng
from ._lib import RobirError

TINY_NUMBER = 1e-6


def _weights_of(VisModel):
    w = getattr(VisModel, "_robir_vis_weights", None)
    if w is None:
        if not hasattr(VisModel, "vis_layer"):
            raise RobirError("robir_b200 needs a VisNetwork-like VisModel (with .vis_layer); got %r" % type(VisModel))
        w = ops.VisWeights(VisModel.vis_layer)
        try:
            VisModel._robir_vis_weights = w
        except Exception:
            pass
    return w


def _batch_min(sg_range):
    """The specular sampler's cone scale is a minimum over the WHOLE batch (model/sg_render.py:220-222): under
    dist.STRONG_SHARDING the ranks hold slices of one batch, so the minimum is all-reduced; its gradient stays with the
    rank that owns the minimum."""
    if not rdist._strong():
        return sg_range
    g = rdist.allreduce_min_scalar(sg_range.detach().clone())
    return torch.where(sg_range <= g, sg_range, g)


def _norm_axis(x):
    return x / (torch.norm(x, dim=-1, keepdim=True) + TINY_NUMBER)


def sample_diffuse_dirs(lgtSGLobes, lgtSGLambdas, nsamp, dev, thr=1.0):
    """Sample directions / weights of get_diffuse_visibility (model/sg_render.py:122-147): they depend on the light SGs
    and the random draws only, so the fixed-capacity forward draws them while the per-point networks run."""
    M = lgtSGLobes.shape[0]
    sharp = torch.clamp(lgtSGLambdas[:, 0], min=1e-4)
    sg_range = torch.clamp(sharp.min(), max=thr).reshape(1)
    u_theta = rng.rand((M, nsamp), dev)
    u_phi = rng.rand((M, nsamp), dev)
    return ops.sample_dirs(lgtSGLobes, lgtSGLobes, sharp, lgtSGLambdas[:, 0], sg_range, u_theta, u_phi, True)


def light_lobes(lgtSGs):
    """(lobes [M,3], lambdas [M,1]) of the shared light SGs as render_with_sg normalises them (sg_render.py:364-366)."""
    return lgtSGs[:, :3] / (torch.norm(lgtSGs[:, :3], dim=-1, keepdim=True) + TINY_NUMBER), torch.abs(lgtSGs[:, 3:4])


def get_diffuse_visibility(points, normals, VisModel, lgtSGLobes, lgtSGLambdas, nsamp=8, testing=False, thr=1.0,
                           bounding=False, argmax_vis=False, presampled=None):
    """[n,3], [n,3], VisModel, lobes [M,3], lambdas [M,1] -> vis [M, n].  presampled: (dirs, w[, tabB]) drawn earlier
    with sample_diffuse_dirs on the same lobes."""
    if bounding or argmax_vis:
        raise RobirError("bounding / argmax_vis variants are not on the accelerated path")
    M = lgtSGLobes.shape[0]
    tabB = None
    if presampled is None:
        dirs, w = sample_diffuse_dirs(lgtSGLobes, lgtSGLambdas, nsamp, points.device, thr)
    else:
        dirs, w = presampled[0], presampled[1]
        tabB = presampled[2] if len(presampled) > 2 else None
    need_grad = torch.is_grad_enabled() and not testing and (dirs.requires_grad or w.requires_grad)
    if testing:
        dirs_q, w_q = dirs.detach(), w   # the reference runs VisModel under no_grad but keeps the weights' graph
    else:
        dirs_q, w_q = dirs, w
    lv = ops.diffuse_vis(points.detach(), normals.detach(), dirs_q, w_q, M, nsamp, _weights_of(VisModel), need_grad,
                         tabB)
    return lv.permute(1, 0)


def _spec_frame(normals, viewdirs):
    ndv = torch.clamp(torch.sum(normals * viewdirs, dim=-1, keepdim=True), min=0.)
    return -viewdirs + 2 * ndv * normals


def get_specular_visibility(points, normals, viewdirs, VisModel, lgtSGLobes, lgtSGLambdas, nsamp=24, multi_view=False,
                            testing=False, inv=False, argmax_vis=False, valid=None):
    """[n,3] x3, VisModel, lobes [n,3], lambdas [n,1] -> vis [n].  valid (static-shape mode): rows excluded from the
    batch-global sharpness minimum (model/sg_render.py:220-222)."""
    if multi_view or argmax_vis:
        raise RobirError("multi_view / argmax_vis variants are not on the accelerated path")
    n = points.shape[0]
    dev = points.device
    ref_dir = _spec_frame(normals, viewdirs)
    sharp = torch.clip(lgtSGLambdas[:, 0], min=0.1, max=50)
    sharp_for_min = sharp if valid is None else torch.where(valid, sharp, torch.full_like(sharp, float("inf")))
    sg_range = _batch_min(torch.clamp(sharp_for_min.min(), max=1).reshape(1))
    u_theta = rng.rand((n, nsamp), dev)
    u_phi = rng.rand((n, nsamp), dev)
    dirs, w = ops.sample_dirs(ref_dir, lgtSGLobes, sharp, sharp, sg_range, u_theta, u_phi, False)
    need_grad = torch.is_grad_enabled() and not testing and (dirs.requires_grad or w.requires_grad)
    dirs_q = dirs.detach() if testing else dirs
    return ops.spec_vis(points.detach(), normals.detach(), dirs_q, w, nsamp, inv, testing, _weights_of(VisModel),
                        need_grad)


def _spec_warp(normal, viewdirs, roughness):
    """Per-point warped BRDF lobe / sharpness (sg_render.py:417-428), needed as sampling input."""
    inv_r4 = 2. / (roughness * roughness * roughness * roughness)
    vdl = torch.clamp(torch.sum(normal * viewdirs, dim=-1, keepdim=True), min=0.)
    wl = 2 * vdl * normal - viewdirs
    wl = wl / (torch.norm(wl, dim=-1, keepdim=True) + TINY_NUMBER)
    return wl, inv_r4 / (4 * vdl + TINY_NUMBER)


def render_with_all_sg(points, normal, viewdirs, lgtSGs, specular_reflectance, roughness, diffuse_albedo,
                       indir_integral=None, indir_lgtSGs=None, VisModel=None, fun_spec=False, lin_diff=False,
                       testing=False, metallic=None, diffuse_vis=None, prefit=False, argmax_vis=False, valid=None,
                       diffuse_presampled=None):
    """model/sg_render.py:304-337 for the PBR-stage configuration (fun_spec=False, metallic=None) and, with
    diffuse_vis [n*M] / prefit, the CESR-stage one (sg_render.py:389-407)."""
    if fun_spec or metallic is not None or argmax_vis or viewdirs.dim() != 2:
        raise RobirError("render_with_all_sg: fun_spec / metallic / argmax_vis / multi-view variants are not on the "
                         "accelerated path")
    if lgtSGs.dim() != 2:
        raise RobirError("render_with_all_sg expects the shared light SGs as [M,7]")
    with ops.point_table_scope():
        return _render_with_all_sg(points, normal, viewdirs, lgtSGs, specular_reflectance, roughness, diffuse_albedo,
                                   indir_integral, indir_lgtSGs, VisModel, lin_diff, testing, valid, diffuse_presampled,
                                   diffuse_vis, prefit)


def kl_divergence(x, mu=0.05, valid=None):
    """utils/utils.py:14-17 (the batch mean spans all ranks under dist.STRONG_SHARDING).  valid [n] bool (fixed-capacity
    mode): the batch mean runs over those rows only."""
    if valid is not None:
        rho_hat = torch.where(valid[:, None], x, torch.zeros_like(x)).sum(0) / valid.sum().clamp(min=1)
    else:
        rho_hat = rdist.batch_mean_rows(x)
    rho = torch.full_like(rho_hat, mu)
    return torch.mean(rho * torch.log(rho / (rho_hat + 1e-4)) + (1 - rho) * torch.log((1 - rho) / (1 - rho_hat + 1e-4)))


def _render_with_all_sg(points, normal, viewdirs, lgtSGs, specular_reflectance, roughness, diffuse_albedo,
                        indir_integral, indir_lgtSGs, VisModel, lin_diff, testing, valid, diffuse_presampled=None,
                        diffuse_vis=None, prefit=False):
    n = normal.shape[0]
    M = lgtSGs.shape[0]
    viewdirs = viewdirs.detach()
    # ---- direct light visibility per lobe (sg_render.py:364-366, 388-391): 32 samples per lobe, 8 when a learned
    # per-lobe visibility (the CESR shadow_net) stands in for it and the MLP average only supervises it
    lobes, lambdas = light_lobes(lgtSGs)
    light_vis = get_diffuse_visibility(points, normal.detach(), VisModel, lobes, lambdas,
                                       nsamp=32 if diffuse_vis is None else 8,
                                       testing=testing, presampled=diffuse_presampled).permute(1, 0)
    supervise = torch.zeros((), device=points.device)
    if diffuse_vis is not None:
        light_vis_gt, light_vis = light_vis, diffuse_vis.reshape(-1, M)
        if prefit == "warmup":                                                          # sg_render.py:397-399
            supervise = kl_divergence((light_vis_gt.detach() - light_vis).abs(), 0.01, valid) * 0.1
            light_vis = light_vis_gt
        else:                                                                           # :400-403
            supervise = kl_divergence((light_vis_gt - light_vis).abs(), 0.01, valid) * \
                (0.2 if prefit == "project" else 1.0)
    # ---- BRDF-lobe visibility, direct then indirect (draw order of SURVEY.md A.4)
    bv_ind = None
    normal_grad = normal.requires_grad and torch.is_grad_enabled() and not testing
    if indir_lgtSGs is not None and indir_integral is None:
        raise RobirError("render_with_all_sg: indirect SGs need indir_integral (PBR-stage configuration)")
    if normal_grad:
        # the shading normal is being trained (CESR after iteration 1000 renders with normal_net's output,
        # train_cesr.py:508; render_with_sg never detaches it, sg_render.py:369-371): the reflection frame and the
        # warped lobe stay in torch so that autograd carries d/d normal, the sample directions / weights and the SG
        # render differentiate on the device (robir_sample_dirs_bwd, robir_sg_render_bwd with g_normal)
        wl, wlam = _spec_warp(normal, viewdirs, roughness)
        bv_dir = get_specular_visibility(points, normal, viewdirs, VisModel, wl, wlam, nsamp=8, testing=testing,
                                         inv=False, valid=valid)
        if indir_lgtSGs is not None:
            bv_ind = get_specular_visibility(points, normal, viewdirs, VisModel, wl, wlam, nsamp=8, testing=testing,
                                             inv=True, valid=valid)
    elif indir_lgtSGs is not None:
        # both get_specular_visibility calls (inv = False / True) share their sampling inputs: one prep kernel, one
        # launch chain over 2n "points" (rows [0, n) direct, [n, 2n) indirect)
        S = 8
        dev = points.device
        ref, wl2, sharp, sg_range = ops.spec_prep(normal, viewdirs, roughness, valid)
        sg_range = _batch_min(sg_range)
        u_theta, u_phi = rng.rand_pairs(n, S, dev)           # theta, phi (direct), theta, phi (indirect)
        dirs, w = ops.sample_dirs(ref, wl2, sharp, sharp, sg_range, u_theta, u_phi, False)
        need_grad = torch.is_grad_enabled() and not testing and (dirs.requires_grad or w.requires_grad)
        bv = ops.spec_vis(points.detach(), normal.detach(), dirs.detach() if testing else dirs, w, S, None, testing,
                          _weights_of(VisModel), need_grad)
        bv_dir, bv_ind = bv[:n], bv[n:]
    else:
        wl, wlam = _spec_warp(normal, viewdirs, roughness)
        bv_dir = get_specular_visibility(points, normal, viewdirs, VisModel, wl, wlam, nsamp=8, testing=testing,
                                         inv=False, valid=valid)
    outs = ops.sg_render(normal if normal_grad else normal.detach(), viewdirs, roughness, diffuse_albedo,
                         specular_reflectance.reshape(1), lgtSGs,
                         indir_lgtSGs, light_vis.contiguous(), bv_dir, bv_ind, indir_integral, lin_diff)
    sg_rgb, sg_spec, sg_diff, vis_shadow, ind_rgb, ind_spec, ind_diff = outs
    return {'sg_rgb': sg_rgb, 'sg_specular_rgb': sg_spec, 'sg_diffuse_rgb': sg_diff, 'vis_shadow': vis_shadow,
            'supervise': supervise, 'indir_rgb': ind_rgb,
            'indir_diffuse_rgb': ind_diff, 'indir_specular_rgb': ind_spec}
