"""``install(model)``: re-bind the hot path on a LIVE reference ``IDRNetwork`` (built by the reference's own code)
without editing reference files.  Seams used:

  1. ``model.ray_tracer`` / ``model.octree_ray_tracer``  -> robir_b200.tracing.OctreeTracing (same call signature,
     ``generate(sdf_fn, tex_sampler)`` discovered via hasattr by the runners, training/train_pbr.py:403-407), or
     robir_b200.sphere_tracing.RayTracing when the model was built with use_octree=False;
  2. ``model.implicit_network.forward / gradient``       -> fused CUDA value / normal kernel reading the module's own
     weight_g / weight_v / bias parameters (state_dict untouched);
  3. ``model.sg_render.render_with_all_sg`` (imported inside the runners' get_sg_render at call time,
     training/train_pbr.py:350) and the module-level name in model.implicit_differentiable_renderer -> ours;
  4. ``model.envmap_material_network.forward`` / ``model.indirect_illum_network.forward`` (rows a6 / a7) -> the fused
     MLP chains of robir_b200.networks, bound as methods onto the reference's own module objects: parameters, buffers
     and state_dict keys stay the reference's (sg_envmap_material.py:40-275, implicit_differentiable_renderer.py:170-222);
  5. ``torch.rand/randn(...).cuda()`` draws stay where the reference makes them (CPU generator, same order, SURVEY.md
     A.4), so seeds reproduce.
"""
import sys
import types

import torch

from . import networks, ops, sg_render, tracing
from ._lib import RobirError


_saved_module_attrs = []


def uninstall_modules():
    """Undo the module-level re-bindings of install() (the per-model attribute changes stay with the model object)."""
    while _saved_module_attrs:
        mod, name, old = _saved_module_attrs.pop()
        setattr(mod, name, old)


def _patch(mod, name, new):
    _saved_module_attrs.append((mod, name, getattr(mod, name)))
    setattr(mod, name, new)


_AE_METHODS = ("_chain", "_wants_grad", "encode_points", "forward_points", "brdf_points")


def _adopt_sparse_ae(ae):
    """Give a reference SparseAE (sg_envmap_material.py:40-118) the fused evaluation methods of networks.SparseAE.
    Its own attributes are what those methods read: brdf_encoder_layer / brdf_decoder_layer (nn.Sequential of Linear +
    LeakyReLU(0.2)), smooth_on_latent, out_act, lc_act, latent_dim and the CESR dropout vector ``var`` (:72,99)."""
    enc = [m for m in ae.brdf_encoder_layer if isinstance(m, torch.nn.Linear)]
    dec = [m for m in ae.brdf_decoder_layer if isinstance(m, torch.nn.Linear)]
    if [l.out_features for l in enc] != [512, 512, 512, 512, ae.latent_dim] or \
            [l.out_features for l in dec[:-1]] != [128, 128]:
        raise RobirError("install(): SparseAE layout differs from the shipped configs (512x4 -> latent -> 128x2)")
    for name in _AE_METHODS:
        setattr(ae, name, types.MethodType(getattr(networks.SparseAE, name), ae))


def install(model, patch_modules=True, stage="PBR"):
    """stage: "PBR" / "CESR" (the indirect-illumination weights are frozen: train_pbr.py:104-105, only the input
    gradient is propagated through them) or "Vis" (train_visibility.py:305-313 steps them: weight gradients on)."""
    if not hasattr(model, "visibility_network") or not hasattr(model, "implicit_network"):
        raise RobirError("install() expects a reference IDRNetwork")
    net = model.implicit_network
    for name in ("ray_tracer", "octree_ray_tracer"):
        old = getattr(model, name, None)
        if old is None:
            continue
        if hasattr(old, "max_iter"):
            new = tracing.OctreeTracing(max_iter=old.max_iter).bind(net)
        else:                             # use_octree=False: the IDR sphere tracer (model/ray_tracing.py, row a3)
            from .sphere_tracing import RayTracing
            new = RayTracing(object_bounding_sphere=old.object_bounding_sphere, sdf_threshold=old.sdf_threshold,
                             line_search_step=old.line_search_step, line_step_iters=old.line_step_iters,
                             sphere_tracing_iters=old.sphere_tracing_iters, n_steps=old.n_steps,
                             n_rootfind_steps=old.n_secant_steps).bind(net)
            new.train(old.training)
        setattr(model, name, new)
    sdfw = ops.SdfWeights(net.neus_model.sdf_network)

    def forward(self, points, compute_grad=False):
        if points.numel() == 0:
            return torch.ones_like(points)
        sdf, _, feat = ops.sdf_eval(sdfw, points, want_feat=True)
        return torch.cat([sdf[:, None], feat], -1)

    def gradient(self, x):
        if x.numel() == 0:
            return torch.ones_like(x)
        return ops.sdf_eval(sdfw, x, want_grad=True)[1].unsqueeze(1)

    net.forward = types.MethodType(forward, net)
    net.gradient = types.MethodType(gradient, net)
    net.sdf = types.MethodType(lambda self, p: ops.sdf_eval(sdfw, p)[0], net)
    net.sdf_and_normal = types.MethodType(lambda self, p: ops.sdf_eval(sdfw, p, want_grad=True)[:2], net)
    # row a13: the secondary-ray radiance (neus_model.py:828-884) and the plain VisNetwork logits (:225-258)
    net._w = sdfw
    for name in ("neus_forward", "borrow_color", "batch_borrow_color"):
        setattr(net, name, types.MethodType(getattr(networks.ImplicitNetworkMy, name), net))
    vis = model.visibility_network
    vis.forward = types.MethodType(networks.VisNetwork.forward, vis)
    # ---- rows a6 / a7: per-point networks on the fused chains (same parameters, reference-owned)
    mat = model.envmap_material_network
    if getattr(mat, "brdf_embed_fn", None) is None or mat.brdf_encoder_layer.brdf_encoder_layer[0].in_features != 63:
        raise RobirError("install(): envmap_material_network must use multires = 10 (confs_sg/*.conf)")
    for ae in (mat.brdf_encoder_layer, mat.spec_brdf_encoder_layer, mat.normal_decoder_layer):
        _adopt_sparse_ae(ae)
    mat.forward = types.MethodType(networks.EnvmapMaterialNetwork.forward, mat)
    ind = model.indirect_illum_network
    if ind.lobe_layer[0].in_features != 64:
        raise RobirError("install(): indirect_illum_network must use multires = 10 with the hdr-shift input")
    _adopt_sparse_ae(ind.integral_layer)
    ind._lobe_chain = None
    ind.train_weights = stage == "Vis"
    ind.forward = types.MethodType(networks.IndirctIllumNetwork.forward, ind)
    if patch_modules:
        for modname in ("model.sg_render", "model.implicit_differentiable_renderer"):
            mod = sys.modules.get(modname)
            if mod is not None:
                _patch(mod, "render_with_all_sg", sg_render.render_with_all_sg)
                if modname == "model.sg_render":
                    _patch(mod, "get_diffuse_visibility", sg_render.get_diffuse_visibility)
                    _patch(mod, "get_specular_visibility", sg_render.get_specular_visibility)
    return model
