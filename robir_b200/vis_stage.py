"""Vis-stage secondary trace: IDRNetwork.trace_radiance
(model/implicit_differentiable_renderer.py:566-650).  Secondary rays are walked by the same cooperative octree kernel
(max_iter = 32 mode), the borrowed radiance at their hit points is the fused SDF value/normal/feature kernel + the fused
colour MLP, and the visibility logits (with gradients, the VisNetwork is trained in this stage) come from the fused MLP
chain."""
import math

import torch

from . import rng


def trace_radiance(model, input, nsamp=16, test_dir=None):
    points, hdr_shift, pm = input["points"], input["hdr_shift"], input["network_object_mask"]
    dev = points.device
    N = points.shape[0]
    trace_rad = torch.zeros(N, nsamp, 3, device=dev)
    idx = pm.nonzero()[:, 0]
    sec_o = points[idx].clone()
    n = sec_o.shape[0]
    sample_dirs = torch.zeros(n, nsamp, 3, device=dev)
    gt_vis = torch.zeros(N, nsamp, 1, dtype=torch.bool, device=dev)
    pred_vis = torch.zeros(N, nsamp, 2, device=dev)
    indir_mask = torch.zeros_like(gt_vis)
    gt_integral = torch.zeros_like(points)
    if n > 0:
        normals = input["normals"].detach()[idx][:, None, :]
        normals = normals / torch.clamp(torch.norm(normals, dim=-1, keepdim=True), 1e-4)
        if test_dir is not None:
            sample_dirs = test_dir[None, None].expand(sample_dirs.shape)
        else:
            u = rng.rand((n * nsamp,), dev) * 2 - 1
            t = rng.rand((n * nsamp,), dev) * math.pi * 2
            s = (1 - u ** 2) ** 0.5
            sample_dirs = torch.stack([s * torch.cos(t), s * torch.sin(t), u], -1).view(n, nsamp, 3)
        back = (normals * sample_dirs).sum(-1) < 0
        with torch.no_grad():
            sec_pts, sec_mask, _ = model.octree_ray_tracer(sdf=model.implicit_network.sdf,
                                                           cam_loc=sec_o + normals[:, 0] * 0.005, object_mask=None,
                                                           ray_directions=sample_dirs.contiguous())
        hit = sec_mask.nonzero()[:, 0]
        if hit.numel() > 0:
            hp = sec_pts[hit]
            hv = -sample_dirs.reshape(-1, 3)[hit]
            rad = torch.zeros_like(sec_pts)
            col = model.implicit_network.batch_borrow_color(hp, hv)
            shift = hdr_shift[idx][:, None, :].expand(-1, nsamp, 1).reshape(-1, 1)[hit]
            rad[hit] = model.gamma.hdr_shift.ldr2hdr(col ** 2.2, shift)
            rad = rad.reshape(n, nsamp, 3)
            rad = torch.where(back[..., None], torch.zeros_like(rad), rad)
            trace_rad[idx] = rad
        in_p = sec_o.unsqueeze(1).expand(-1, nsamp, 3)
        logits = model.visibility_network(in_p.reshape(-1, 3), sample_dirs.reshape(-1, 3))
        pred_vis = pred_vis.index_copy(0, idx, logits.reshape(-1, nsamp, 2))
        gv = sec_mask.reshape(n, nsamp, 1)
        gt_vis[idx] = gv
        indir_mask[idx] = ~back[..., None] & gv
        cos_dot = trace_rad[idx] * torch.relu((normals * sample_dirs).sum(-1, keepdim=True))
        hemi = (~back).sum(-1)[..., None]
        gt_integral[idx] = cos_dot.sum(-2) / torch.clamp(hemi, 1e-4)
    return {'trace_radiance': trace_rad, 'sample_dirs': sample_dirs, 'gt_vis': gt_vis, 'pred_vis': pred_vis,
            'indir_mask': indir_mask[..., 0], 'gt_integral': gt_integral}
