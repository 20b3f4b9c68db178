"""PBR-stage loss, the consumer of the hot path's outputs (model/loss.py:7-125 InvLoss, training/train_pbr.py:313-346
pbr_step / white_loss): ``InvLoss`` is the reference-shaped torch form; ``fused_pbr_loss`` / ``pbr_step_loss`` run the
same terms (value + every input gradient) on one kernel (csrc/loss.cu)."""
import ctypes

import torch
import torch.nn as nn

from . import dist as rdist
from .networks import positional_encoding


class InvLoss(nn.Module):
    def __init__(self, idr_rgb_weight=1.0, eikonal_weight=0.1, mask_weight=100.0, alpha=50.0, sg_rgb_weight=1.0,
                 kl_weight=1.0, latent_smooth_weight=1.0, brdf_multires=10, loss_type='L1'):
        super().__init__()
        if brdf_multires != 10:
            raise ValueError("InvLoss: the hit points are re-encoded with 10 frequencies (shipped configs, "
                             "model/loss.py:20,86); brdf_multires = %r is not supported" % (brdf_multires,))
        self.sg_rgb_weight, self.kl_weight, self.latent_smooth_weight = sg_rgb_weight, kl_weight, latent_smooth_weight
        self.l2 = loss_type == 'L2'
        self.static_shapes = False

    @staticmethod
    def kl_divergence(rho, latent):
        rho_hat = rdist.batch_mean_rows(torch.sigmoid(latent))
        rho = torch.full_like(rho_hat, rho)
        return torch.mean(rho * torch.log(rho / (rho_hat + 1e-4))
                          + (1 - rho) * torch.log((1 - rho) / (1 - rho_hat + 1e-4)))

    def forward(self, model_outputs, ground_truth, mat_model=None, train_idr=False, train_spec=False, hdr_fn=None):
        rgb_gt = ground_truth['rgb'].to(model_outputs['sg_rgb'].device)
        nm = model_outputs['network_object_mask'] & model_outputs['object_mask']
        pred = model_outputs['sg_rgb'] + model_outputs['indir_rgb']
        pred = hdr_fn(pred) if hdr_fn is not None else pred / (pred + 1)
        diff = pred - rgb_gt.reshape(-1, 3)
        per = (diff * diff if self.l2 else diff.abs()) * nm[:, None]
        # per-ray terms are normalised by the counts of the WHOLE batch (all ranks under dist.STRONG_SHARDING)
        sg_rgb_loss = per.sum() / rdist.global_count(model_outputs['object_mask'].shape[0], per)
        smooth = rdist.global_mean((model_outputs['diffuse_albedo'] - model_outputs['random_xi_diffuse_albedo']).abs()) + \
            rdist.global_mean((model_outputs['roughness'][..., 0] - model_outputs['random_xi_roughness'][..., 0]).abs()) \
            * 0.2
        enc = mat_model.spec_brdf_encoder_layer if train_spec else mat_model.brdf_encoder_layer
        sm = model_outputs['surface_mask']

        def latent_of(pts):
            # the reference re-encodes the hit points (loss.py:86-88); the forward already did exactly that, so the
            # fused path reuses its latent (identical values, gradients add up in the same graph node)
            z = getattr(mat_model, "_last_spec_latent", None) if train_spec else None
            if z is not None and z.shape[0] == pts.shape[0] and z.requires_grad == torch.is_grad_enabled():
                return z
            if pts.is_cuda and hasattr(enc, "encode_points"):
                return enc.encode_points(pts, "pe10")
            return enc.encode(positional_encoding(pts, 10))

        if self.static_shapes:
            if rdist.STRONG_SHARDING:
                raise RuntimeError("InvLoss: the fixed-capacity mode reduces over the local batch only; strong sharding "
                                   "of one batch (dist.STRONG_SHARDING) needs the dynamic-shape mode")
            # same statistics without data-dependent shapes (CUDA-graph capturable): masked batch mean of the latent
            hit = model_outputs['network_object_mask']
            pts = torch.where(hit[:, None], model_outputs['points'], torch.zeros_like(model_outputs['points']))
            z = latent_of(pts)
            if z is getattr(mat_model, "_last_spec_latent", None) and \
                    getattr(mat_model, "_last_latent_valid", None) is not None:
                hit = mat_model._last_latent_valid      # the cached latent lives in the renderer's compacted row order
            lat = torch.sigmoid(z)
            rho_hat = (lat * hit[:, None]).sum(0) / hit.sum().clamp(min=1)
            rho = torch.full_like(rho_hat, 0.05)
            kl = torch.mean(rho * torch.log(rho / (rho_hat + 1e-4))
                            + (1 - rho) * torch.log((1 - rho) / (1 - rho_hat + 1e-4)))
            normal_loss = torch.zeros((), device=pts.device)     # reported only, never part of the loss (loss.py:114-123)
        else:
            pts = model_outputs['points'][model_outputs['network_object_mask']]
            kl = self.kl_divergence(0.05, latent_of(pts))
            normal_loss = ((model_outputs['normal_map'][sm] - model_outputs['normals'][sm]) ** 2).mean()
        return {'sg_rgb_loss': sg_rgb_loss, 'kl_loss': self.kl_weight * kl,
                'latent_smooth_loss': self.latent_smooth_weight * smooth, 'normal_loss': normal_loss,
                'loss': self.sg_rgb_weight * sg_rgb_loss}


def white_loss(lgtSGs):
    lgt = torch.abs(lgtSGs[..., -3:])
    mu = lgt.norm(dim=-1, keepdim=True) + 1e-4
    return (lgt / mu).var(-1).mean() * 0.01


class _FusedPBRLoss(torch.autograd.Function):
    """loss = w_rgb * L(hdr2ldr(sg_rgb + indir_rgb), gt) + kl + 0.1 * smooth + white in one launch (csrc/loss.cu);
    the kernel also writes every input gradient, so backward is a scaling by the upstream gradient."""

    @staticmethod
    def forward(ctx, sg_rgb, indir_rgb, adapt_illum, albedo, albedo_r, rough, rough_r, z, lgt, gt, mask, z_valid, cfg,
                hit=None, order=None):
        from . import _lib
        from ._lib import LossParams, check, lib, ptr, stream
        w_rgb, w_kl, w_smooth, l2 = cfg
        N, n_lat, M = sg_rgb.shape[0], z.shape[0], lgt.shape[0]
        dev = sg_rgb.device

        def rows(t, cols):          # [N, >= cols] view with unit column stride -> (tensor, row stride)
            t = t.detach()
            if t.dtype != torch.float32 or t.stride(-1) != 1 or t.dim() != 2:
                t = t.float().reshape(t.shape[0], -1).contiguous()
            return t, t.stride(0)

        ts = {}
        p = LossParams()
        p.N, p.n_lat, p.M, p.l2 = N, n_lat, M, int(l2)
        for name, t, ld in (("sg_rgb", sg_rgb, "ld_sg"), ("indir_rgb", indir_rgb, "ld_ind"), ("albedo", albedo, "ld_alb"),
                            ("albedo_r", albedo_r, "ld_albr"), ("rough", rough, "ld_r"), ("rough_r", rough_r, "ld_rr")):
            v, st = rows(t, 3)
            ts[name] = v
            setattr(p, name, ctypes.c_void_p(v.data_ptr()))       # views: data_ptr() already includes the offset
            setattr(p, ld, st)
        gt_c = gt.detach().to(dev).reshape(-1, 3).float().contiguous()
        mask_c = mask.detach().to(torch.uint8).contiguous()
        z_c = z.detach().float().contiguous()
        zv = z_valid.detach().to(torch.uint8).contiguous() if z_valid is not None else None
        lgt_c = lgt.detach().float().contiguous()
        a_c = adapt_illum.detach().float().reshape(1).contiguous()
        p.gt, p.mask, p.adapt_illum, p.z, p.z_valid, p.lgt = ptr(gt_c), ptr(mask_c), ptr(a_c), ptr(z_c), ptr(zv), ptr(lgt_c)
        if order is not None:
            hit_c = hit.detach().to(torch.uint8).contiguous()
            order_c = order.detach().to(torch.int64).contiguous()
            p.hit, p.order = ptr(hit_c), ptr(order_c)
        p.w_rgb, p.w_kl, p.w_smooth, p.rho = float(w_rgb), float(w_kl), float(w_smooth), 0.05
        losses = torch.empty(5, device=dev)
        # every gradient lives in one flat buffer: the backward scales it with a single launch
        sizes = dict(g_pred=N * 3, g_adapt=1, g_albedo=N * 3, g_albedo_r=N * 3, g_rough=N, g_rough_r=N, g_z=n_lat * 32,
                     g_lgt=M * 7)
        flat = torch.empty(sum(-(-v // 4) * 4 for v in sizes.values()), device=dev)
        g, off = {}, 0
        for k, v in sizes.items():
            g[k] = (off, v)
            off += -(-v // 4) * 4
        p.losses = ptr(losses)
        for k, (o, v) in g.items():
            setattr(p, k, ctypes.c_void_p(flat.data_ptr() + 4 * o))
        check(lib().robir_pbr_loss(ctypes.byref(p), stream()))
        ctx.save_for_backward(flat)
        ctx.layout = (g, N, n_lat, M)
        ctx.shapes = (tuple(adapt_illum.shape), tuple(rough.shape), tuple(rough_r.shape))
        ctx.mark_non_differentiable(losses)
        return losses[0], losses

    @staticmethod
    def backward(ctx, g_loss, _g_all):
        flat, = ctx.saved_tensors
        g, N, n_lat, M = ctx.layout
        sa, sr, srr = ctx.shapes
        scaled = flat * g_loss
        view = lambda k, *shape: scaled[g[k][0]:g[k][0] + g[k][1]].view(*shape)
        g_pred = view("g_pred", N, 3)

        def col0(k, shape):         # roughness arrives as [N, 1] or [N, 3] (expanded): the loss reads column 0 only
            if shape[1] == 1:
                return view(k, N, 1)
            out = torch.zeros(shape, device=flat.device)
            out[:, 0] = view(k, N)
            return out
        return (g_pred, g_pred, view("g_adapt", 1).reshape(sa), view("g_albedo", N, 3), view("g_albedo_r", N, 3),
                col0("g_rough", sr), col0("g_rough_r", srr), view("g_z", n_lat, 32), view("g_lgt", M, 7),
                None, None, None, None, None, None)


def fused_pbr_loss(model, loss_fn, model_outputs, ground_truth):
    """pbr_step_loss on the fused kernel; None when its preconditions do not hold (CPU tensors, no cached latent)."""
    mat = model.envmap_material_network
    z = getattr(mat, "_last_spec_latent", None)
    o = model_outputs
    if z is None or not o['sg_rgb'].is_cuda or z.shape[1] != 32 or not z.requires_grad == torch.is_grad_enabled():
        return None
    if loss_fn.static_shapes:
        z_valid = getattr(mat, "_last_latent_valid", None)
        if z_valid is None or z.shape[0] != z_valid.shape[0]:
            return None
    else:
        z_valid = None
        if z.shape[0] != int(o['network_object_mask'].sum()):
            return None
    nm = o['network_object_mask'] & o['object_mask']
    cfg = (loss_fn.sg_rgb_weight, loss_fn.kl_weight * 1.0, loss_fn.latent_smooth_weight * 0.1, loss_fn.l2)
    c = getattr(model, "_static_compact", None) if loss_fn.static_shapes else None
    if c is not None and c["ret_sg_rgb"] is o['sg_rgb']:
        # fixed-capacity forward: feed the kernel the compacted tensors (row i = ray order[i]); same value, and the
        # backward goes straight into the render / material graph without the un-permute
        loss, parts = _FusedPBRLoss.apply(c['sg_rgb'], c['indir_rgb'], model.gamma.hdr_shift.adapt_illum,
                                          c['diffuse_albedo'], c['random_xi_diffuse_albedo'], c['roughness'],
                                          c['random_xi_roughness'], z, mat.lgtSGs, ground_truth['rgb'], nm, z_valid, cfg,
                                          o['network_object_mask'], c['order'])
    else:
        loss, parts = _FusedPBRLoss.apply(o['sg_rgb'], o['indir_rgb'], model.gamma.hdr_shift.adapt_illum,
                                          o['diffuse_albedo'], o['random_xi_diffuse_albedo'], o['roughness'],
                                          o['random_xi_roughness'], z, mat.lgtSGs, ground_truth['rgb'], nm, z_valid, cfg)
    if loss_fn.static_shapes:
        normal_loss = torch.zeros((), device=loss.device)
    else:
        with torch.no_grad():
            sm = o['surface_mask']
            normal_loss = ((o['normal_map'][sm] - o['normals'][sm]) ** 2).mean()
    return loss, {'sg_rgb_loss': parts[1], 'kl_loss': loss_fn.kl_weight * parts[2],
                  'latent_smooth_loss': loss_fn.latent_smooth_weight * parts[3], 'normal_loss': normal_loss,
                  'loss': loss_fn.sg_rgb_weight * parts[1]}


FUSED_LOSS = True


def pbr_step_loss(model, loss_fn, model_outputs, ground_truth, train_spec=True):
    if FUSED_LOSS and train_spec and not rdist.STRONG_SHARDING:      # the fused kernel reduces over the local batch only
        fused = fused_pbr_loss(model, loss_fn, model_outputs, ground_truth)
        if fused is not None:
            return fused
    out = loss_fn(model_outputs, ground_truth, mat_model=model.envmap_material_network, train_idr=False,
                  train_spec=train_spec, hdr_fn=model.gamma.hdr_shift.hdr2ldr)
    loss = out['loss'] + out['kl_loss'] * 1.0 + out['latent_smooth_loss'] * 0.1
    return loss + rdist.param_only(white_loss(model.envmap_material_network.lgtSGs)), out


def query_indir_illum(lgtSGs, sample_dirs):
    """Radiance of the per-point indirect SG mixture along sample_dirs (model/loss.py:128-141).
    lgtSGs [n,Mi,7], sample_dirs [n,S,3] -> [n,S,3]."""
    lobes = lgtSGs[:, None, :, :3] / torch.norm(lgtSGs[:, None, :, :3], dim=-1, keepdim=True)
    lam, mu = lgtSGs[:, None, :, 3:4], lgtSGs[:, None, :, -3:]
    cos = torch.sum(sample_dirs[:, :, None, :] * lobes, dim=-1, keepdim=True)
    return torch.sum(mu * torch.exp(lam * (cos - 1.)), dim=2)


class IllumLoss(nn.Module):
    """Vis-stage losses (model/loss.py:144-179): radiance L1 of the indirect SGs against the traced radiance plus the
    integral term, and the visibility cross-entropy."""

    def __init__(self, loss_type='L1'):
        super().__init__()
        self.rgb_loss = nn.L1Loss(reduction='mean') if loss_type == 'L1' else nn.MSELoss(reduction='mean')

    def forward(self, model_outputs, trace_outputs, anneal_t):
        indir_mask = trace_outputs["indir_mask"]
        pm = model_outputs["network_object_mask"]
        lgtSGs = model_outputs["indirect_sgs"][pm]
        gt_radiance = trace_outputs['trace_radiance'][indir_mask] + anneal_t
        pred_radiance = query_indir_illum(lgtSGs, trace_outputs['sample_dirs'])
        radiance_loss = self.rgb_loss(gt_radiance, pred_radiance[indir_mask[pm]])
        radiance_loss = radiance_loss + self.rgb_loss(trace_outputs['gt_integral'][pm],
                                                      model_outputs['indir_integral'][pm])
        gt_vis = (~trace_outputs['gt_vis'][pm]).long().reshape(-1)
        pred_vis = trace_outputs['pred_vis'][pm].reshape(-1, 2)
        return radiance_loss, nn.functional.cross_entropy(pred_vis, gt_vis)
