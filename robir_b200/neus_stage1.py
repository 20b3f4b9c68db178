"""Stage-1 NeuS volume renderer, evaluation (forward) path on the CUDA kernels.

``render_neus(net, rays_o, rays_d, near, far, ...)`` mirrors neus/volume_render/sdf_render.py:236-348 with the shipped
configuration (``n_outside = 0``, ``lindisp = False``): 64 uniform depths (+ the optional per-ray jitter), 4 rounds of
importance sampling from the section-wise alpha of the current SDF samples at inv_s = 64 * 2^i (``up_sample`` :38-82,
``sample_pdf`` :5-35, ``cat_z_vals`` :85-99), then ``render_core`` (:141-233) on the 128 merged depths: SDF value +
normal + features, colour network, NeuS alpha, transmittance, composited colour / weights / depth / accumulation and
the relaxed Eikonal statistic.  ``net`` is ``robir_b200.networks.ImplicitNetworkMy`` (the stage-1 checkpoint is its
``neus_model.*`` sub-tree) or a reference ``ImplicitNetworkMy`` after ``robir_b200.install``.  All coordinates are
NeuS coordinates.

Kernels: ``robir_neus_upsample`` / ``robir_neus_merge`` / ``robir_neus_midpoints`` / ``robir_neus_composite``
(csrc/neus.cu, one warp per ray) around ``robir_sdf_eval`` (value + forward-mode normal + 256 features in one pass) and
the fused colour chain.  The training step of stage 1 differentiates THROUGH the SDF normal (Eikonal term and the
normal-conditioned colour network: a second-order graph); that backward is not built -- this module is the renderer the
stage-1 evaluation / mesh-extraction paths call, and it refuses to run under autograd.
"""
import torch

from . import ops
from ._lib import RobirError, check, f32, lib, ptr, stream

RADIUS = 2.0          # NeuSModel.radius() (model/neus_model.py:743-744)


def _sdf(net, pts):
    return ops.sdf_eval(net._w, pts, in_scale=1.0, sdf_scale=1.0, feat_scale=1.0)[0]


def up_sample(rays_o, rays_d, z_vals, sdf, n_importance, inv_s, radius=RADIUS):
    B, n = z_vals.shape
    new_z = torch.empty(B, n_importance, device=z_vals.device)
    check(lib().robir_neus_upsample(B, n, n_importance, ptr(rays_o), ptr(rays_d), ptr(z_vals), ptr(sdf), float(inv_s),
                                    float(radius), ptr(new_z), stream()))
    return new_z


def cat_z_vals(net, rays_o, rays_d, z_vals, new_z, sdf, last):
    B, n = z_vals.shape
    m = new_z.shape[1]
    out_z = torch.empty(B, n + m, device=z_vals.device)
    if last:
        check(lib().robir_neus_merge(B, n, m, ptr(z_vals), None, ptr(new_z), None, ptr(out_z), None, stream()))
        return out_z, sdf
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * new_z[..., :, None]).reshape(-1, 3).contiguous()
    new_sdf = _sdf(net, pts).reshape(B, m).contiguous()
    out_sdf = torch.empty(B, n + m, device=z_vals.device)
    check(lib().robir_neus_merge(B, n, m, ptr(z_vals), ptr(sdf), ptr(new_z), ptr(new_sdf), ptr(out_z), ptr(out_sdf),
                                 stream()))
    return out_z, out_sdf


def render_core(net, rays_o, rays_d, z_vals, sample_dist, near, far, white_bkgd=True, cos_anneal_ratio=0.0):
    B, n = z_vals.shape
    dev = z_vals.device
    mid_z = torch.empty(B, n, device=dev)
    pts = torch.empty(B * n, 3, device=dev)
    check(lib().robir_neus_midpoints(B, n, float(sample_dist), ptr(rays_o), ptr(rays_d), ptr(z_vals), ptr(mid_z),
                                     ptr(pts), stream()))
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3).contiguous()
    color, sdf, grad = net.neus_forward(pts, dirs, return_grad=True)
    inv_s = float(torch.exp(net.neus_model.deviation_network.variance.detach() * 10.0).clip(1e-6, 1e6))
    rgb, weights = torch.empty(B, 3, device=dev), torch.empty(B, n, device=dev)
    acc, dist = torch.empty(B, device=dev), torch.empty(B, device=dev)
    eik = torch.zeros(2, device=dev)
    sdf_c, grad_c, color_c = f32(sdf).reshape(-1), f32(grad), f32(color)     # named: copies must outlive the launch
    check(lib().robir_neus_composite(B, n, float(sample_dist), inv_s, float(cos_anneal_ratio), RADIUS, int(white_bkgd),
                                     ptr(rays_o), ptr(rays_d), ptr(z_vals), ptr(sdf_c), ptr(grad_c),
                                     ptr(color_c), ptr(near), ptr(far), ptr(rgb), ptr(weights), ptr(acc), ptr(dist),
                                     ptr(eik), stream()))
    return dict(color=rgb, weights=weights, mid_z_vals=mid_z, acc=acc, dist=dist,
                gradient_error=eik[0] / (eik[1] + 1e-5), sdf=sdf, gradients=grad.reshape(B, n, 3))


def render_neus(net, rays_o, rays_d, near, far, t_rand=None, cos_anneal_ratio=1.0, n_samples=64, n_importance=64,
                up_sample_steps=4, white_bkgd=True):
    """rays_o / rays_d [B,3], near / far [B,1] -> dict(rgb [B,3], dist [B], acc [B], sim_or_grad [], weights [B,128],
    means [B,128]).  t_rand [B,1]: the per-ray jitter of the training-time sampler (None = evaluation)."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in net.parameters()):
        raise RobirError("robir_b200.neus_stage1.render_neus is the evaluation path (the second-order stage-1 training "
                         "backward is not built): call it under torch.no_grad()")
    rays_o, rays_d = f32(rays_o), f32(rays_d)
    near, far = f32(near).reshape(-1, 1), f32(far).reshape(-1, 1)
    if not rays_o.is_cuda:
        raise RobirError("render_neus needs CUDA tensors (there is no CPU path)")
    B = rays_o.shape[0]
    if n_samples + n_importance > 256 or n_importance % max(up_sample_steps, 1):
        raise RobirError("render_neus: at most 256 depths per ray, n_importance divisible by up_sample_steps")
    sample_dist = 2.0 / n_samples
    z_vals = near + (far - near) * torch.linspace(0.0, 1.0, n_samples, device=rays_o.device)[None, :]
    if t_rand is not None:
        z_vals = z_vals + (f32(t_rand) - 0.5) * 2.0 / n_samples
    z_vals = z_vals.contiguous()
    if n_importance > 0:
        pts = (rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]).reshape(-1, 3).contiguous()
        sdf = _sdf(net, pts).reshape(B, n_samples).contiguous()
        for i in range(up_sample_steps):
            new_z = up_sample(rays_o, rays_d, z_vals, sdf, n_importance // up_sample_steps, 64 * 2 ** i)
            z_vals, sdf = cat_z_vals(net, rays_o, rays_d, z_vals, new_z, sdf, last=(i + 1 == up_sample_steps))
    fine = render_core(net, rays_o, rays_d, z_vals, sample_dist, near.reshape(-1).contiguous(),
                       far.reshape(-1).contiguous(), white_bkgd, cos_anneal_ratio)
    return dict(rgb=fine["color"], dist=fine["dist"], acc=fine["acc"], sim_or_grad=fine["gradient_error"],
                weights=fine["weights"], means=fine["mid_z_vals"])
