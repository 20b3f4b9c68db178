"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  (see oracle/robir_oracle.py header for the import rules)

SURVEY.md section 8f rank 4: the stage-1 NeuS volume renderer (geometry pre-training), restated as pure functions over
the same flat state dict as the rest of the oracle (the stage-1 checkpoint is the ``implicit_network.neus_model.*``
sub-dict of the stage-2 state dict).  Reference: neus/volume_render/sdf_render.py -- sample_pdf :5-35 (det=True),
up_sample :38-82, cat_z_vals :85-99, render_core :141-233, render_neus :236-348 with the shipped configuration
``render_neus.n_outside = 0`` (every neus/config/*.gin; no background model).  All coordinates are NeuS coordinates.

Parity status: PINNED against the unmodified reference file executed in the build container
(tests/test_oracle_vs_reference.py::test_neus_stage1_render_matches_reference; the file is loaded with stub ``misc``
modules and driven with the reference's stage-2 copy of NeuSModel, which implements the same ISDF interface).  No CUDA
product exists for this row yet: this module is the checker a later round builds against.

Random draws: ``t_rand`` [batch, 1] uniform (perturb > 0) is passed in explicitly.
"""
import torch
import torch.nn.functional as F

import robir_oracle as O

VAR_KEY = "implicit_network.neus_model.deviation_network.variance"
RADIUS = 2.0   # NeuSModel.radius() of the stage-2 copy (model/neus_model.py:743-744); stage 1 reads it from the model


def sdf_and_feat(sd, pts):
    out = O.sdf_network(sd, pts)
    return out[:, :1], out[:, 1:]


def sdf_gradient(sd, pts, create_graph):
    """SDFNetwork.gradient (model/neus_model.py:424-438): d sdf / d x, with a graph when training (the Eikonal term
    back-propagates through it)."""
    with torch.enable_grad():
        x = pts if pts.requires_grad else pts.detach().clone().requires_grad_(True)
        y = O.sdf_network(sd, x)[:, :1]
        g = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph, retain_graph=create_graph)[0]
    return g


def inv_s_of(sd):
    """SingleVarianceNetwork (model/neus_model.py:644-650) as used at sdf_render.py:169."""
    return torch.exp(sd[VAR_KEY] * 10.0).clip(1e-6, 1e6)


def sample_pdf_det(bins, weights, n_samples):
    """sdf_render.py:5-35 with det=True (the only mode up_sample uses)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0. + 0.5 / n_samples, 1. - 0.5 / n_samples, steps=n_samples)
    u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)


def up_sample(rays_o, rays_d, z_vals, sdf, n_importance, inv_s, sphere_radius):
    """sdf_render.py:38-82: importance samples from the section-wise alpha of the current SDF samples at a fixed inv_s."""
    batch, n = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
    radius = torch.linalg.norm(pts, ord=2, dim=-1)
    inside = (radius[:, :-1] < sphere_radius) | (radius[:, 1:] < sphere_radius)
    sdf = sdf.reshape(batch, n)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z_vals[:, :-1], z_vals[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-5)
    prev_cos = torch.cat([torch.zeros(batch, 1), cos_val[:, :-1]], dim=-1)
    cos_val = torch.minimum(prev_cos, cos_val).clip(-1e3, 0.0) * inside
    dist = next_z - prev_z
    prev_cdf = torch.sigmoid((mid_sdf - cos_val * dist * 0.5) * inv_s)
    next_cdf = torch.sigmoid((mid_sdf + cos_val * dist * 0.5) * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    weights = alpha * torch.cumprod(torch.cat([torch.ones(batch, 1), 1. - alpha + 1e-7], -1), -1)[:, :-1]
    return sample_pdf_det(z_vals, weights, n_importance).detach()


def cat_z_vals(sd, rays_o, rays_d, z_vals, new_z_vals, sdf, last):
    """sdf_render.py:85-99: merge + sort the new depths; evaluate the SDF at them unless this is the last round."""
    batch, n = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z_vals[..., :, None]
    z_all, index = torch.sort(torch.cat([z_vals, new_z_vals], dim=-1), dim=-1)
    if not last:
        new_sdf = O.sdf_network(sd, pts.reshape(-1, 3))[:, :1].reshape(batch, -1)
        sdf = torch.gather(torch.cat([sdf, new_sdf], dim=-1), 1, index)
    return z_all, sdf


def render_core(sd, rays_o, rays_d, z_vals, sample_dist, background_rgb=None, cos_anneal_ratio=0.0, training=True):
    """sdf_render.py:141-233 without a background model (n_outside = 0)."""
    batch, n = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
    mid_z = z_vals + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(batch, n, 3).reshape(-1, 3)
    if training:
        pts = pts.detach().clone().requires_grad_(True)      # the reference's gradient() flags the points in place
    sdf, feat = sdf_and_feat(sd, pts)
    gradients = sdf_gradient(sd, pts, create_graph=training)
    color = O.color_network(sd, pts, gradients, dirs, feat).reshape(batch, n, 3)
    inv_s = inv_s_of(sd).expand(batch * n, 1)
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    next_sdf = sdf + iter_cos * dists.reshape(-1, 1) * 0.5
    prev_sdf = sdf - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf, next_cdf = torch.sigmoid(prev_sdf * inv_s), torch.sigmoid(next_sdf * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).reshape(batch, n).clip(0.0, 1.0)
    pts_norm = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).reshape(batch, n)
    inside = (pts_norm < RADIUS).float().detach()
    relax_inside = (pts_norm < RADIUS * 1.2).float().detach()
    alpha = alpha * inside
    weights = alpha * torch.cumprod(torch.cat([torch.ones(batch, 1), 1. - alpha + 1e-7], -1), -1)[:, :-1]
    rgb = (color * weights[:, :, None]).sum(dim=1)
    if background_rgb is not None:
        rgb = rgb + background_rgb * (1.0 - weights.sum(dim=-1, keepdim=True))
    g3 = gradients.reshape(batch, n, 3)
    eik = (torch.linalg.norm(g3, ord=2, dim=-1) - 1.0) ** 2
    eik = (relax_inside * eik).sum() / (relax_inside.sum() + 1e-5)
    return dict(color=rgb, sdf=sdf, dists=dists, gradients=g3, s_val=1.0 / inv_s, mid_z_vals=mid_z, weights=weights,
                cdf=prev_cdf.reshape(batch, n), gradient_error=eik, inside_sphere=inside)


def render_neus(sd, rays_o, rays_d, near, far, t_rand, cos_anneal_ratio, n_samples=64, n_importance=64,
                up_sample_steps=4, white_bkgd=True, training=True):
    """sdf_render.py:236-348 with n_outside = 0, lindisp = False.  t_rand [batch,1] uniform, or None for perturb = 0 /
    is_eval."""
    batch = rays_o.shape[0]
    sample_dist = 2.0 / n_samples
    z_vals = near + (far - near) * torch.linspace(0.0, 1.0, n_samples)[None, :]
    if t_rand is not None:
        z_vals = z_vals + (t_rand - 0.5) * 2.0 / n_samples
    if n_importance > 0:
        with torch.no_grad():
            pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
            sdf = O.sdf_network(sd, pts.reshape(-1, 3))[:, :1].reshape(batch, n_samples)
            for i in range(up_sample_steps):
                new_z = up_sample(rays_o, rays_d, z_vals, sdf, n_importance // up_sample_steps, 64 * 2 ** i, RADIUS)
                z_vals, sdf = cat_z_vals(sd, rays_o, rays_d, z_vals, new_z, sdf, last=(i + 1 == up_sample_steps))
    fine = render_core(sd, rays_o, rays_d, z_vals, sample_dist,
                       background_rgb=torch.ones(1, 3) if white_bkgd else None, cos_anneal_ratio=cos_anneal_ratio,
                       training=training)
    weights = fine["weights"]
    acc = weights.sum(dim=-1)
    distance = (weights[..., :128] * fine["mid_z_vals"]).sum(dim=-1) / acc
    distance = torch.clip(torch.nan_to_num(distance, torch.inf), near.squeeze(), far.squeeze())
    return dict(rgb=fine["color"], dist=distance, acc=acc, sim_or_grad=fine["gradient_error"], weights=weights,
                means=fine["mid_z_vals"])


def stage1_loss(ret, mask, pixels, eikonal_weight=0.1):
    """neus/optimization/trainer.py:136-175 with the shipped regulariser set (neus/config/blender.gin: eikonal only):
    masked MSE of the composited colour + eikonal_weight * the relaxed Eikonal term that render_core already reduced
    (regular.py:41-43 just scales it).  mask = rays.lossmult [B,1]."""
    mse = (mask * (ret["rgb"] - pixels[..., :3]) ** 2).sum() / (mask.sum() + 1e-5)
    return mse + eikonal_weight * ret["sim_or_grad"].sum(), dict(mse=mse)
