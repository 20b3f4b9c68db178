"""Model facade with the reference's surface: ``IDRNetwork.forward(input, trainstage, fun_spec, lin_diff, train_spec)``
and ``IDRNetwork.trace_radiance(input, nsamp, test_dir)`` return the same dict keys / shapes / dtypes
(model/implicit_differentiable_renderer.py:261-650), ``get_sg_render`` is the re-bindable hook (:400-409, :499-529) and
``pbr_get_sg_render`` is the PBR runner's version of it (training/train_pbr.py:348-396)."""
import types

import numpy as np
import torch
import torch.nn as nn

from . import networks, ops, rng, sg_render, tracing
from ._lib import RobirError


def _cfg(conf, key, default=None):
    if conf is None:
        return default
    node = conf
    try:
        for part in key.split("."):
            node = node[part]
        return node
    except (KeyError, TypeError):
        return default


class IDRNetwork(nn.Module):
    def __init__(self, conf=None):
        super().__init__()
        if not _cfg(conf, "use_neus", True) or not _cfg(conf, "use_octree", True):
            if not _cfg(conf, "use_neus", True):
                raise RobirError("only use_neus=True models are supported (all shipped configs)")
        rt = dict(_cfg(conf, "ray_tracer", {}) or {})
        self.feature_vector_size = int(_cfg(conf, "feature_vector_size", 256))
        self.octree_ray_tracer = tracing.OctreeTracing(**rt, max_iter=32)
        if _cfg(conf, "use_octree", True):
            self.ray_tracer = tracing.OctreeTracing(**rt)
        else:
            from .sphere_tracing import RayTracing
            self.ray_tracer = RayTracing(**rt)
        self.object_bounding_sphere = float(rt.get("object_bounding_sphere", 1.0))
        self.implicit_network = networks.ImplicitNetworkMy()
        self.indirect_illum_network = networks.IndirctIllumNetwork(
            num_lgt_sgs=int(_cfg(conf, "indirect_illum_network.num_lgt_sgs", 24)))
        self.visibility_network = networks.VisNetwork()
        self.envmap_material_network = networks.EnvmapMaterialNetwork(
            num_lgt_sgs=int(_cfg(conf, "envmap_material_network.num_lgt_sgs", 128)),
            upper_hemi=bool(_cfg(conf, "envmap_material_network.upper_hemi", False)),
            specular_albedo=float(_cfg(conf, "envmap_material_network.specular_albedo", 0.05)))
        self.gamma = networks.GammaCorrect(float(_cfg(conf, "gamma", 1.0)), int(_cfg(conf, "hdr_mode", 0)))
        # PBR-runner state read by pbr_get_sg_render (training/train_pbr.py:414-415, 131-159)
        self.no_normal = True
        self.is_training = True
        # static_shapes=True: no hit compaction (no host sync, no data-dependent shapes) -- every ray is carried through
        # the per-hit stages with a validity mask, which makes the whole training step capturable in one CUDA graph
        # (robir_b200.graph.GraphedPBRStep).  Hit rows get bit-identical results; random draws have shape [N, .].
        self.static_shapes = False

    # ------------------------------------------------------------------------------------------------------------------
    def generate(self):
        """What every runner does before its loop (train_pbr.py:403-407)."""
        sdf_fn = lambda x: self.implicit_network(x)[:, 0]
        if isinstance(self.ray_tracer, tracing.OctreeTracing):
            self.ray_tracer.generate(sdf_fn, None, implicit_network=self.implicit_network)
            self.octree_ray_tracer.sdf_octree = self.ray_tracer.sdf_octree   # same tree, different max_iter
        else:   # use_octree=False: the IDR sphere tracer has nothing to build (no generate(), like the reference's)
            self.ray_tracer.bind(self.implicit_network)
            self.octree_ray_tracer.generate(sdf_fn, None, implicit_network=self.implicit_network)

    def prepack(self):
        """Touch the packed weight copies of every fused chain that already exists (they are created lazily by the
        first forward); a no-op for copies that are up to date."""
        for mod in self.modules():
            for v in list(mod.__dict__.values()):
                if isinstance(v, ops.MlpChain):
                    v.packed()
        return None

    def _trace(self, tracer, cam_loc, object_mask, ray_dirs):
        return tracer(sdf=self.implicit_network.sdf, cam_loc=cam_loc, object_mask=object_mask,
                      ray_directions=ray_dirs)

    def forward(self, input, trainstage='IDR', fun_spec=False, lin_diff=False, train_spec=False):
        if fun_spec:
            raise RobirError("fun_spec=True is not on the accelerated path")
        if self.static_shapes and trainstage != 'Illum' and "intrinsics" in input and 'hdr_shift' in input:
            hook = None
            if getattr(self.get_sg_render, "__func__", None) is not IDRNetwork.get_sg_render:
                hook = getattr(self.get_sg_render, "__self__", None)
                if not hasattr(hook, "get_sg_render_static"):
                    raise RobirError("the fixed-capacity forward (static_shapes = True) implements the PBR runner's hook "
                                     "and robir_b200.cesr.ClusteredAlbedoHook; any other re-bound get_sg_render needs "
                                     "static_shapes = False")
            return self._forward_static(input, lin_diff=lin_diff, train_spec=train_spec, hook=hook)
        if "intrinsics" in input:
            object_mask = input["object_mask"].reshape(-1)
            ray_dirs, cam_loc = ops.camera_rays(input["uv"], input["pose"], input["intrinsics"])
            batch_size, num_pixels, _ = ray_dirs.shape
            with torch.no_grad():
                points, network_object_mask, dists = self._trace(self.ray_tracer, cam_loc, object_mask, ray_dirs)
            shard = input.get("shard")
            if shard is not None:
                # strong sharding of ONE batch over ranks (dist.STRONG_SHARDING): `uv` / `object_mask` / `hdr_shift` hold
                # the WHOLE batch and every rank walks all of it -- the octree walk's per-iteration sample count depends
                # on the number of live rays of the whole batch (utils/octree.py:542-548), and the walk is latency-bound,
                # so the replicated trace costs what the sharded one would -- then shades rays [lo, hi) only
                lo, hi = int(shard[0]), int(shard[1])
                object_mask, network_object_mask, dists = object_mask[lo:hi], network_object_mask[lo:hi], dists[lo:hi]
                ray_dirs, num_pixels = ray_dirs[:, lo:hi], hi - lo
                input = dict(input)
                for k in ("hdr_shift", "albedo_ratio"):
                    if k in input and input[k] is not None and input[k].shape[0] == points.shape[0]:
                        input[k] = input[k][lo:hi]
        else:
            cam_loc = input["points"].reshape(-1, 3)
            ray_dirs = input["dirs"].reshape(-1, 1, 3)
            object_mask = input["object_mask"].reshape(-1) if "object_mask" in input else \
                torch.ones_like(ray_dirs[..., 0, 0], dtype=torch.bool)
            batch_size, num_pixels, _ = ray_dirs.shape
            points = torch.zeros_like(cam_loc)
            network_object_mask = torch.zeros_like(object_mask)
            dists = torch.zeros_like(cam_loc[..., 0])
            with torch.no_grad():
                p, m, d = self._trace(self.ray_tracer, cam_loc[object_mask], object_mask[object_mask],
                                      ray_dirs[object_mask])
                points[object_mask], network_object_mask[object_mask], dists[object_mask] = p, m, d
        points = (cam_loc.unsqueeze(1) + dists.reshape(batch_size, num_pixels, 1) * ray_dirs).reshape(-1, 3)
        sdf_output = self.implicit_network.sdf(points)[:, None]
        ray_dirs = ray_dirs.reshape(-1, 3)
        ret = {'points': points, 'sdf_output': sdf_output, 'network_object_mask': network_object_mask,
               'object_mask': object_mask, 'ray_dirs': ray_dirs}
        surface_mask = network_object_mask
        hit_idx = surface_mask.nonzero()[:, 0]          # one host sync per call, like the reference's boolean indexing
        n_hit = hit_idx.shape[0]
        total = points.shape[0]
        dev = points.device
        hit_points = points[hit_idx]
        n_ind = self.indirect_illum_network.num_lgt_sgs
        indirect_sgs = torch.ones(total, n_ind, 7, device=dev)
        indirect_sgs[:, :, -3:] = 0
        indirect_integral = torch.ones(total, 3, device=dev)
        hit_sgs = hit_int = None
        if 'hdr_shift' in input:
            if n_hit > 0:
                hit_sgs, hit_int = self.indirect_illum_network(hit_points, input['hdr_shift'][hit_idx])
                indirect_sgs = indirect_sgs.index_copy(0, hit_idx, hit_sgs)
                indirect_integral = indirect_integral.index_copy(0, hit_idx, hit_int)
            ret['hdr_shift'] = input['hdr_shift']
        if trainstage == 'Illum':
            ret.update({'indirect_sgs': indirect_sgs, 'indir_integral': indirect_integral})
            normals = torch.ones_like(points)
            if n_hit > 0:
                mat = self.envmap_material_network(hit_points, train_spec=False, train_norm=True)
                normals = normals.index_copy(0, hit_idx, mat["sg_normal_map"])
            ret['normals'] = normals
            return ret

        ones3 = lambda: torch.ones(total, 3, device=dev)
        ones1 = lambda: torch.ones(total, 1, device=dev)
        buf = {k: ones3() for k in ('sg_rgb', 'indir_rgb', 'sg_diffuse_rgb', 'sg_specular_rgb', 'indir_diffuse_rgb',
                                    'indir_specular_rgb', 'normals', 'diffuse_albedo', 'roughness', 'normal_map',
                                    'vis_shadow', 'random_xi_diffuse_albedo', 'random_xi_roughness')}
        buf['metallic'] = ones1()
        buf['random_xi_metallic'] = ones1()
        gradient_error = torch.tensor(0.0, device=dev)
        if n_hit > 0:
            if hit_sgs is None:
                hit_sgs, hit_int = indirect_sgs[hit_idx], indirect_integral[hit_idx]
            r = self.get_sg_render(hit_points, -ray_dirs[hit_idx], hit_sgs,
                                   albedo_ratio=input.get('albedo_ratio'), fun_spec=fun_spec, lin_diff=lin_diff,
                                   train_spec=train_spec, indir_integral=hit_int,
                                   tex_uv=None, hdr_shift=input['hdr_shift'][hit_idx] if 'hdr_shift' in input else None)
            for k in buf:
                if k not in r:
                    continue
                v = r[k]
                if k in ('roughness', 'random_xi_roughness'):
                    v = v.expand(-1, 3)
                buf[k] = buf[k].index_copy(0, hit_idx, v)
            if 'gradient_error' in r:                   # the CESR hook's supervise term (:446-447)
                gradient_error = gradient_error + r['gradient_error']
        ret.update({'final_t': ones1(), 'gradient_error': gradient_error, 'acc': ones1(),
                    'bg_rgb': ones3(), 'surface_mask': surface_mask})
        ret.update(buf)
        return ret

    def trace_rays(self, uv, pose, intrinsics, object_mask):
        """Camera rays + surface trace of one batch, outside autograd: (ray_dirs [1,N,3], cam_loc [1,3], mask [N] bool,
        dists [N]).  Nothing here depends on the parameters the PBR / CESR stages train, which is what lets
        graph.GraphedPBRStep(pipeline_trace=True) walk the NEXT batch while the current step is in its backward."""
        ray_dirs, cam_loc = ops.camera_rays(uv, pose, intrinsics)
        with torch.no_grad():
            _, mask, dists = self._trace(self.ray_tracer, cam_loc, object_mask.reshape(-1), ray_dirs)
        return ray_dirs, cam_loc, mask, dists

    def _forward_static(self, input, lin_diff=False, train_spec=False, hook=None):
        object_mask = input["object_mask"].reshape(-1)
        traced = input.get("traced")
        if traced is not None:                      # (ray_dirs, cam_loc, mask, dists) of trace_rays for exactly this batch
            ray_dirs, cam_loc, mask, dists = traced
            batch_size, num_pixels, _ = ray_dirs.shape
            self.prepack()
        else:
            ray_dirs, cam_loc = ops.camera_rays(input["uv"], input["pose"], input["intrinsics"])
            batch_size, num_pixels, _ = ray_dirs.shape
            def trace():
                with torch.no_grad():
                    return self._trace(self.ray_tracer, cam_loc, object_mask, ray_dirs)
            # packed copies of the weights that are being trained are rebuilt while the tracer walks the octree
            (_, mask, dists), _ = ops.fork_join([trace, self.prepack], tag="trace")
        points = (cam_loc.unsqueeze(1) + dists.reshape(batch_size, num_pixels, 1) * ray_dirs).reshape(-1, 3)
        ray_dirs = ray_dirs.reshape(-1, 3)
        total, dev = points.shape[0], points.device
        # hit rays are compacted to the front of the fixed-capacity batch (device-side permutation, no host sync); the
        # per-point networks then evaluate n_act rows instead of N (ops.active_rows), like the reference's boolean
        # indexing (implicit_differentiable_renderer.py:341-347) but capturable in a CUDA graph
        pos, order, n_act, valid, pts, view = ops.compact_hits(mask, points, ray_dirs)   # inactive rows: finite inputs
        hdr = input['hdr_shift'].index_select(0, order)
        # the indirect-illumination net, the material net, the SDF normal and the reported sdf_output are independent
        # given the hit points: parallel branches (random draws keep the reference order: indirect, BRDF latent,
        # normal input)
        def act(fn):
            def run():
                with ops.active_rows(n_act):
                    return fn()
            return run
        # sdf_output is only reported (no consumer inside the step): it runs on its own low-priority stream and joins
        # at the end of the forward instead of holding up the visibility kernels
        main = torch.cuda.current_stream()
        if getattr(self, "_aux_stream", None) is None or self._aux_stream.device != main.device:
            self._aux_stream = torch.cuda.Stream(priority=0)
        aux = self._aux_stream
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            sdf_output = self.implicit_network.sdf(points)[:, None]
        def diffuse_samples():
            # light-lobe sample directions and their layer-0 table only need the light SGs and the random draws: made
            # while the networks run (last in the list: its draws follow the networks' draws, as in the reference)
            mnet = self.envmap_material_network
            lgt = mnet.lgtSGs
            if mnet.upper_hemi:
                lgt = torch.cat((lgt[..., :1], torch.abs(lgt[..., 1:2]), lgt[..., 2:]), dim=-1)
            lobes, lambdas = sg_render.light_lobes(lgt)
            # 32 samples per lobe; 8 when the CESR shadow_net stands in for the average (sg_render.py:388-391)
            dirs, w = sg_render.sample_diffuse_dirs(lobes, lambdas, 32 if hook is None else 8, dev)
            if not ops.vis_engine_terms():
                return dirs, w
            W = sg_render._weights_of(self.visibility_network).get()
            return dirs, w, ops.pe_linear(dirs.detach(), W["Wt0d"], None)
        (sgs, integ), mat, nrm, presampled = ops.fork_join([
            act(lambda: self.indirect_illum_network(pts, hdr)),
            act(lambda: self.envmap_material_network(pts, train_spec=train_spec)),
            act(lambda: self.get_idr_render(pts, None, normal_only=True)),
            diffuse_samples], tag="nets")
        self.envmap_material_network._last_latent_valid = valid
        ret = {'points': points, 'sdf_output': sdf_output, 'network_object_mask': mask, 'object_mask': object_mask,
               'ray_dirs': ray_dirs, 'hdr_shift': input['hdr_shift']}
        if hook is None:
            r = pbr_get_sg_render(self, pts, view, sgs, lin_diff=lin_diff, train_spec=train_spec, indir_integral=integ,
                                  valid=valid, precomputed=(nrm, mat), diffuse_presampled=presampled)
        else:
            r = hook.get_sg_render_static(pts, view, sgs, lin_diff=lin_diff, train_spec=train_spec, indir_integral=integ,
                                          valid=valid, n_act=n_act, precomputed=(nrm, mat), diffuse_presampled=presampled)
        # back to ray order, 1.0 for the rays that missed (implicit_differentiable_renderer.py:365-385): one gather over
        # the column-concatenated outputs instead of one per tensor
        keys = ('sg_rgb', 'indir_rgb', 'sg_diffuse_rgb', 'sg_specular_rgb', 'indir_diffuse_rgb', 'indir_specular_rgb',
                'normals', 'diffuse_albedo', 'normal_map', 'vis_shadow', 'random_xi_diffuse_albedo', 'metallic',
                'random_xi_metallic', 'roughness', 'random_xi_roughness')
        # only what the PBR loss differentiates keeps its graph (the others would drag zero gradients through the
        # normal auto-encoder's backward; the reference's gradient contract has them at zero)
        with_grad = ('sg_rgb', 'indir_rgb', 'diffuse_albedo', 'roughness', 'random_xi_diffuse_albedo',
                     'random_xi_roughness', 'sg_diffuse_rgb', 'sg_specular_rgb', 'indir_diffuse_rgb',
                     'indir_specular_rgb')
        wide = torch.cat([r[k] if k in with_grad else r[k].detach() for k in keys], 1)
        wide = torch.where(mask[:, None], wide.index_select(0, pos), torch.ones((), device=dev))
        c = 0
        for k in keys:
            w = r[k].shape[1]
            ret[k] = wide[:, c:c + w]
            c += w
        ret['roughness'] = ret['roughness'].expand(-1, 3)
        ret['random_xi_roughness'] = ret['random_xi_roughness'].expand(-1, 3)
        total, dev = points.shape[0], points.device
        # the same quantities in compacted row order, for robir_b200.loss.pbr_step_loss: its fused kernel reads them
        # directly (gathering gt / masks through `order`), so the backward never walks the un-permute above
        self._static_compact = dict(order=order, ret_sg_rgb=ret['sg_rgb'], sg_rgb=r['sg_rgb'], indir_rgb=r['indir_rgb'],
                                    diffuse_albedo=r['diffuse_albedo'], roughness=r['roughness'],
                                    random_xi_diffuse_albedo=r['random_xi_diffuse_albedo'],
                                    random_xi_roughness=r['random_xi_roughness'])
        main.wait_stream(aux)
        if not torch.cuda.is_current_stream_capturing():
            sdf_output.record_stream(main)
        ret.update({'final_t': torch.ones(total, 1, device=dev),
                    'gradient_error': r['gradient_error'] if 'gradient_error' in r else torch.zeros((), device=dev),
                    'acc': torch.ones(total, 1, device=dev), 'bg_rgb': torch.ones(total, 3, device=dev),
                    'surface_mask': mask})
        return ret

    # ------------------------------------------------------------------------------------------------------------------
    def get_idr_render(self, points, view_dirs=None, normal_only=False):
        if not normal_only:
            raise RobirError("get_idr_render(normal_only=False) is not on the accelerated path")
        return self.implicit_network.gradient(points)[:, 0, :]

    def get_sg_render(self, points, view_dirs, indir_lgtSGs, albedo_ratio=None, fun_spec=False, lin_diff=False,
                      train_spec=False, **kwargs):
        """Default hook = the PBR runner's (the runners re-bind this attribute: train_pbr.py:413)."""
        return pbr_get_sg_render(self, points, view_dirs, indir_lgtSGs, albedo_ratio=albedo_ratio, fun_spec=fun_spec,
                                 lin_diff=lin_diff, train_spec=train_spec, **kwargs)

    # ------------------------------------------------------------------------------------------------------------------
    def trace_radiance(self, input, nsamp=16, test_dir=None):
        from .vis_stage import trace_radiance
        return trace_radiance(self, input, nsamp=nsamp, test_dir=test_dir)


def pbr_get_sg_render(model, points, view_dirs, indir_lgtSGs, albedo_ratio=None, fun_spec=False, lin_diff=False,
                      train_spec=False, indir_integral=None, valid=None, precomputed=None, diffuse_presampled=None,
                      **kwargs):
    """training/train_pbr.py:348-396 (model.no_normal / model.is_training play the runner's attributes).
    valid: optional [n] bool mask of the static-shape mode (rows that are not surface hits are carried along with a
    zero normal, which culls all of their visibility queries)."""
    view_dirs = view_dirs / (torch.norm(view_dirs, dim=-1, keepdim=True) + 1e-6)
    normals = precomputed[0] if precomputed is not None else model.get_idr_render(points, view_dirs, normal_only=True)
    normals = normals / torch.clamp(torch.norm(normals, dim=-1, keepdim=True), 1e-4)
    if valid is not None:
        normals = torch.where(valid[:, None], normals, torch.zeros_like(normals))
    ret = {'normals': normals}
    mat = precomputed[1] if precomputed is not None else model.envmap_material_network(points, train_spec=train_spec)
    indir_integral = indir_integral * 2 * np.pi
    normal_map = mat['sg_normal_map']
    sg = sg_render.render_with_all_sg(points=points.detach(),
                                      normal=normals.detach() if model.no_normal else (
                                          normal_map.detach() if valid is None else
                                          torch.where(valid[:, None], normal_map.detach(), torch.zeros_like(normal_map))),
                                      viewdirs=view_dirs, lgtSGs=mat['sg_lgtSGs'], indir_integral=indir_integral,
                                      specular_reflectance=mat['sg_specular_reflectance'].abs(),
                                      roughness=mat['sg_roughness'], diffuse_albedo=mat['sg_diffuse_albedo'],
                                      indir_lgtSGs=indir_lgtSGs, VisModel=model.visibility_network, fun_spec=False,
                                      lin_diff=False, testing=not model.is_training, metallic=None, valid=valid,
                                      diffuse_presampled=diffuse_presampled)
    ret.update(sg)
    ret.update({'diffuse_albedo': mat['sg_diffuse_albedo'], 'roughness': mat['sg_roughness'],
                'metallic': mat['sg_metallic'], 'normal_map': normal_map,
                'random_xi_roughness': mat['random_xi_roughness'], 'random_xi_metallic': mat['random_xi_metallic'],
                'random_xi_diffuse_albedo': mat['random_xi_diffuse_albedo']})
    return ret
