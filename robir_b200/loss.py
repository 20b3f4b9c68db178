"""PBR-stage loss, the consumer of the hot path's outputs (model/loss.py:7-125 InvLoss, training/train_pbr.py:313-346
pbr_step / white_loss).  Elementwise torch glue; fusing it into the render epilogue is a 'next' row."""
import torch
import torch.nn as nn

from .networks import positional_encoding


class InvLoss(nn.Module):
    def __init__(self, idr_rgb_weight=1.0, eikonal_weight=0.1, mask_weight=100.0, alpha=50.0, sg_rgb_weight=1.0,
                 kl_weight=1.0, latent_smooth_weight=1.0, brdf_multires=10, loss_type='L1'):
        super().__init__()
        self.sg_rgb_weight, self.kl_weight, self.latent_smooth_weight = sg_rgb_weight, kl_weight, latent_smooth_weight
        self.l2 = loss_type == 'L2'
        self.static_shapes = False

    @staticmethod
    def kl_divergence(rho, latent):
        rho_hat = torch.mean(torch.sigmoid(latent), 0)
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
        sg_rgb_loss = per.sum() / float(model_outputs['object_mask'].shape[0])
        smooth = (model_outputs['diffuse_albedo'] - model_outputs['random_xi_diffuse_albedo']).abs().mean() + \
            (model_outputs['roughness'][..., 0] - model_outputs['random_xi_roughness'][..., 0]).abs().mean() * 0.2
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


def pbr_step_loss(model, loss_fn, model_outputs, ground_truth, train_spec=True):
    out = loss_fn(model_outputs, ground_truth, mat_model=model.envmap_material_network, train_idr=False,
                  train_spec=train_spec, hdr_fn=model.gamma.hdr_shift.hdr2ldr)
    loss = out['loss'] + out['kl_loss'] * 1.0 + out['latent_smooth_loss'] * 0.1
    return loss + white_loss(model.envmap_material_network.lgtSGs), out


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
