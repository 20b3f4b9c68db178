"""``install(model)``: re-bind the hot path on a LIVE reference ``IDRNetwork`` (built by the reference's own code)
without editing reference files.  Seams used:

  1. ``model.ray_tracer`` / ``model.octree_ray_tracer``  -> robir_b200.tracing.OctreeTracing (same call signature,
     ``generate(sdf_fn, tex_sampler)`` discovered via hasattr by the runners, training/train_pbr.py:403-407), or
     robir_b200.sphere_tracing.RayTracing when the model was built with use_octree=False;
  2. ``model.implicit_network.forward / gradient``       -> fused CUDA value / normal kernel reading the module's own
     weight_g / weight_v / bias parameters (state_dict untouched);
  3. ``model.sg_render.render_with_all_sg`` (imported inside the runners' get_sg_render at call time,
     training/train_pbr.py:350) and the module-level name in model.implicit_differentiable_renderer -> ours;
  4. ``torch.rand/randn(...).cuda()`` draws stay where the reference makes them (CPU generator), so seeds reproduce.
"""
import sys
import types

import torch

from . import ops, sg_render, tracing
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


def install(model, patch_modules=True):
    if not hasattr(model, "visibility_network") or not hasattr(model, "implicit_network"):
        raise RobirError("install() expects a reference IDRNetwork")
    net = model.implicit_network
    for name in ("ray_tracer", "octree_ray_tracer"):
        old = getattr(model, name, None)
        if old is None:
            continue
        if hasattr(old, "max_iter"):
            new = tracing.OctreeTracing(max_iter=old.max_iter)
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
    if patch_modules:
        for modname in ("model.sg_render", "model.implicit_differentiable_renderer"):
            mod = sys.modules.get(modname)
            if mod is not None:
                _patch(mod, "render_with_all_sg", sg_render.render_with_all_sg)
                if modname == "model.sg_render":
                    _patch(mod, "get_diffuse_visibility", sg_render.get_diffuse_visibility)
                    _patch(mod, "get_specular_visibility", sg_render.get_specular_visibility)
    return model
