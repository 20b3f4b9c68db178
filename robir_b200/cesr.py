"""CESR-stage extras (training/train_cesr.py): the two extra weight-normed MLPs

    shadow_net = SDFNetwork(63 + 128, 2, 512, 8, [4], 0)      per (hit point, light lobe) visibility classifier
    normal_net = SDFNetwork(63, 3, 512, 8, [4], 0)            per hit point normal

and the runner's ``get_sg_render`` hook (train_cesr.py:465-544) that feeds ``diffuse_vis`` / ``prefit`` into
``render_with_all_sg`` and builds the ``supervise`` term, plus the stage's step loss (:387-430).

``WnMLP`` keeps the reference module's state-dict layout (``lin{l}.weight_g | weight_v | bias``), so checkpoints of the
reference's ``*-shadow.pth`` / ``*-normal.pth`` (train_cesr.py:266-276) load unchanged, and the hook equally accepts the
reference's own ``SDFNetwork`` objects (it only reads those three tensors per layer).

Engines for the 512-wide chains (``ops.ENGINE["wn"]``):
  * ``"tc"``    -- the tcgen05 layer engine (csrc/tc_mlp.cu, bf16 hi/lo 3-term products, fp32 accumulation): one launch per
                   layer forward, one per layer for the input-gradient chain, ``robir_mlp_wgrad`` for dW / db;
  * ``"torch"`` -- plain torch (cuBLAS) layers; kept as the cross-check of the engine in the GPU tests.
Both need the CUDA extension to be present for the rest of the hook; there is no CPU path.
"""
import math

import numpy as np
import torch
from torch import nn

from . import dist as rdist
from . import ops, sg_render
from ._lib import RobirError
from .loss import white_loss
from .networks import _WNLinear, positional_encoding


class WnMLP(nn.Module):
    """The reference SDFNetwork with multires = 0 (model/neus_model.py:312-417): ``n_layers + 1`` weight-normed linears,
    the skip layers take cat([h, inputs]) / sqrt(2), softplus(beta=100) between layers; geometric initialisation as at
    :358-376 (multires = 0 branch)."""

    def __init__(self, d_in, d_out, d_hidden=512, n_layers=8, skip_in=(4,), bias=0.5):
        super().__init__()
        dims = [d_in] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        self.d_in = d_in
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if l + 1 in self.skip_in else dims[l + 1]
            lin = _WNLinear(dims[l], out_dim)
            with torch.no_grad():
                if l == self.num_layers - 2:
                    lin.weight_v.normal_(mean=math.sqrt(math.pi) / math.sqrt(dims[l]), std=0.0001)
                    lin.bias.fill_(-bias)
                else:
                    lin.weight_v.normal_(0.0, math.sqrt(2) / math.sqrt(out_dim))
                lin.weight_g.copy_(lin.weight_v.norm(dim=1, keepdim=True))
            setattr(self, "lin%d" % l, lin)

    def forward(self, inputs):
        if inputs.numel() == 0:
            return torch.ones_like(inputs)
        shape = list(inputs.shape[:-1]) + [-1]
        return wn_mlp(self, inputs.reshape(-1, inputs.shape[-1])).reshape(shape)


def _layers(net):
    n_lin = net.num_layers - 1
    return [getattr(net, "lin%d" % l) for l in range(n_lin)], tuple(net.skip_in)


def wn_mlp(net, x, n_active=None):
    """Forward of a WnMLP / reference SDFNetwork(multires=0) on rows x [R, d_in] (the reference evaluates the same rows
    in 1024-row chunks, neus_model.py:397-415).  n_active (device int32 [1], fixed-capacity batches): only the leading
    n_active rows carry data; the tensor-core engine skips the row tiles beyond them (see ops.wn_chain)."""
    lins, skip = _layers(net)
    if getattr(net, "embed_fn_fine", None) is not None or getattr(net, "scale", 1) != 1:
        raise RobirError("wn_mlp: only the multires = 0, scale = 1 form of SDFNetwork (CESR shadow_net / normal_net)")
    # weight-norm: W[o, :] = g[o] v[o, :] / ||v[o, :]||  (tiny; autograd carries dW back to g and v)
    Ws = [l.weight_g * l.weight_v / l.weight_v.norm(dim=1, keepdim=True) for l in lins]
    bs = [l.bias for l in lins]
    if not x.is_cuda:
        raise RobirError("robir_b200.cesr needs CUDA tensors (there is no CPU path)")
    ops.lib()                                   # fail loudly when the extension is missing, whichever engine is selected
    if ops.ENGINE.get("wn", "tc") == "tc" and x.shape[0] >= 128:
        return ops.wn_chain(x, Ws, bs, skip, n_active=n_active)
    return _wn_rows_torch(Ws, bs, skip, x)


def _wn_rows_torch(Ws, bs, skip, x):
    """Library-GEMM form of the chain: the cross-check of the layer engine, and the path of batches below one row tile."""
    h = x
    for l, (W, b) in enumerate(zip(Ws, bs)):
        if l in skip:
            h = torch.cat([h, x], 1) / np.sqrt(2)
        h = nn.functional.linear(h, W, b)
        if l < len(Ws) - 1:
            h = nn.functional.softplus(h, beta=100)
    return h


class ClusteredAlbedoHook:
    """What ``ClusteredAlbedoTrainRunner`` contributes to the hot path: the ``get_sg_render`` hook, the warm-up /
    explore / project schedule (train_cesr.py:546-559) and the step loss (:387-430).  Bind it with
    ``model.get_sg_render = hook.get_sg_render`` (the reference does the same at :588); ``model`` is a
    ``robir_b200.IDRNetwork`` (dynamic-shape forward) or a reference ``IDRNetwork`` after ``robir_b200.install``."""

    def __init__(self, model, shadow_net=None, normal_net=None, white_light=False, explore_iter=1000, proj_iter=0,
                 explore_smooth=0.1, explore_kl=1.0, proj_smooth=0.01, proj_kl=0.01, cur_iter=0, runner=None):
        """runner: optional live ``ClusteredAlbedoTrainRunner``; ``cur_iter`` / ``is_training`` / ``train_spec`` are then
        read from it at every call (the reference's loop advances ``self.cur_iter`` on the runner, train_cesr.py:636)."""
        self.model = model
        self.runner = runner
        self.shadow_embed, in_dim = (lambda x: positional_encoding(x, 10)), 63           # get_embedder(10), :106
        self.shadow_net = shadow_net if shadow_net is not None else WnMLP(in_dim + 128, 2)
        self.normal_net = normal_net if normal_net is not None else WnMLP(in_dim, 3)
        self.white_light = white_light
        self.explore_iter, self.proj_iter = explore_iter, proj_iter
        self.weights = dict(explore=(explore_smooth, explore_kl), project=(proj_smooth, proj_kl))
        self._state = dict(cur_iter=cur_iter, train_spec=True, is_training=True)

    def _get(self, key):
        if self.runner is not None and hasattr(self.runner, key):
            return getattr(self.runner, key)
        return self._state[key]

    cur_iter = property(lambda self: self._get("cur_iter"), lambda self, v: self._state.__setitem__("cur_iter", v))
    train_spec = property(lambda self: self._get("train_spec"), lambda self, v: self._state.__setitem__("train_spec", v))
    is_training = property(lambda self: self._get("is_training"),
                           lambda self, v: self._state.__setitem__("is_training", v))

    @classmethod
    def bind(cls, runner):
        """Re-bind the seam of a live reference runner (what train_cesr.py:588 does with its own method): the hook
        reads the runner's networks, conf and counters; nothing of the runner is modified except
        ``runner.model.get_sg_render``.  Returns the hook."""
        conf = runner.conf
        try:
            argmax_vis = conf.get_bool('train.argmax_vis')          # train_cesr.py:522
        except Exception:
            argmax_vis = False
        if argmax_vis:
            raise RobirError("ClusteredAlbedoHook: train.argmax_vis = True is not on the accelerated path "
                             "(sg_render.py:170,270 hard visibility); the shipped configs leave it off")
        hook = cls(runner.model, runner.shadow_net, runner.normal_net, white_light=runner.white_light,
                   explore_iter=conf.get_int('train.explore_iter'), proj_iter=conf.get_int('train.proj_iter'),
                   explore_smooth=conf.get_float('train.explore_smooth'), explore_kl=conf.get_float('train.explore_kl'),
                   proj_smooth=conf.get_float('train.proj_smooth'), proj_kl=conf.get_float('train.proj_kl'),
                   runner=runner)
        runner.model.get_sg_render = hook.get_sg_render
        return hook

    # ---- schedule (train_cesr.py:546-559)
    def is_explore_step(self):
        if self.cur_iter > 500:
            return self.cur_iter % (self.explore_iter + self.proj_iter) >= self.proj_iter
        return False

    def prefit_option(self):
        if not self.is_explore_step():
            return "warmup" if self.cur_iter <= 500 else "project"
        return "explore"

    def phase_key(self):
        """Everything of the schedule that changes the step's control flow (what a captured CUDA graph bakes in)."""
        return (self.prefit_option(), self.cur_iter > 1000, self.cur_iter > 500, bool(self.is_training),
                bool(self.train_spec))

    def parameters(self):
        return list(self.shadow_net.parameters()) + list(self.normal_net.parameters())

    # ---- the hook (train_cesr.py:465-544)
    def get_sg_render(self, points, view_dirs, indir_lgtSGs, albedo_ratio=None, fun_spec=False, lin_diff=False,
                      train_spec=False, indir_integral=None, **kwargs):
        model = self.model
        if fun_spec:
            raise RobirError("fun_spec=True is not on the accelerated path")
        view_dirs = view_dirs / (torch.norm(view_dirs, dim=-1, keepdim=True) + 1e-6)
        normals = model.get_idr_render(points, view_dirs, normal_only=True)
        normals = normals / torch.clamp(torch.norm(normals, dim=-1, keepdim=True), 1e-4)
        ret = {'normals': normals}
        assert train_spec == self.train_spec
        mat = model.envmap_material_network(points, train_spec=train_spec)
        lgtSGs = mat['sg_lgtSGs']
        M = lgtSGs.shape[0]
        if M != 128:
            raise RobirError("the CESR hook is written for 128 light lobes (train_cesr.py:492-493), got %d" % M)
        indir_integral = indir_integral * 2 * np.pi
        diffuse_albedo, roughness = mat['sg_diffuse_albedo'], mat['sg_roughness']
        normal_map = mat['sg_normal_map'].detach()

        emb = self.shadow_embed(points.detach())                                        # [n, 63]
        with torch.set_grad_enabled(self.is_training and torch.is_grad_enabled()):
            diffuse_vis = shadow_logits(self.shadow_net, emb, M)                        # [n * M, 2]
            normal_new = wn_mlp(self.normal_net, emb)
        normal_new = normal_new / torch.clamp(normal_new.norm(dim=-1, keepdim=True), 1e-4)
        diffuse_vis = torch.softmax(diffuse_vis, -1)[..., 1]
        prefit = self.prefit_option()
        # after iteration 1000 the reference renders with normal_net's normals and lets the render loss train them
        # (:508): render_with_all_sg differentiates with respect to a normal that requires grad
        sg = sg_render.render_with_all_sg(points=points.detach(),
                                          normal=normal_new if self.cur_iter > 1000 else normal_map,
                                          viewdirs=view_dirs, lgtSGs=lgtSGs, indir_integral=indir_integral,
                                          specular_reflectance=mat['sg_specular_reflectance'].abs(),
                                          roughness=roughness, diffuse_albedo=diffuse_albedo,
                                          indir_lgtSGs=indir_lgtSGs, VisModel=model.visibility_network, fun_spec=False,
                                          lin_diff=True, testing=not self.is_training, metallic=None,
                                          diffuse_vis=diffuse_vis, prefit=prefit, argmax_vis=False)
        sg["sg_rgb"] = sg["sg_diffuse_rgb"] * diffuse_albedo / np.pi + sg["sg_specular_rgb"]
        sg["indir_rgb"] = sg["indir_diffuse_rgb"] * diffuse_albedo / np.pi + sg["indir_specular_rgb"]
        supervise = sg['supervise']
        if self.white_light and prefit != "warmup":
            supervise = supervise + rdist.param_only(white_loss(lgtSGs))
        supervise = supervise + rdist.global_mean((normal_map - normal_new) ** 2)
        ret.update(sg)
        ret.update({'diffuse_albedo': diffuse_albedo, 'roughness': roughness, 'metallic': mat['sg_metallic'],
                    'normal_map': normal_new, 'gradient_error': supervise,
                    'random_xi_roughness': mat['random_xi_roughness'],
                    'random_xi_metallic': mat['random_xi_metallic'],
                    'random_xi_diffuse_albedo': mat['random_xi_diffuse_albedo']})
        return ret

    # ---- the same hook for the fixed-capacity forward (IDRNetwork._forward_static; CUDA-graph mode, graph.GraphedPBRStep
    # with hook=): rows are the whole ray batch, hits compacted to the front, `valid` marks them, `n_act` counts them on
    # the device.  Batch statistics (the two supervise means) run over the valid rows; the two extra networks skip the
    # row tiles beyond n_act.
    def get_sg_render_static(self, points, view_dirs, indir_lgtSGs, lin_diff=False, train_spec=False,
                             indir_integral=None, valid=None, n_act=None, precomputed=None, diffuse_presampled=None):
        model = self.model
        view_dirs = view_dirs / (torch.norm(view_dirs, dim=-1, keepdim=True) + 1e-6)
        normals = precomputed[0]
        normals = normals / torch.clamp(torch.norm(normals, dim=-1, keepdim=True), 1e-4)
        normals = torch.where(valid[:, None], normals, torch.zeros_like(normals))
        ret = {'normals': normals}
        assert train_spec == self.train_spec
        mat = precomputed[1]
        lgtSGs = mat['sg_lgtSGs']
        M = lgtSGs.shape[0]
        if M != 128:
            raise RobirError("the CESR hook is written for 128 light lobes (train_cesr.py:492-493), got %d" % M)
        indir_integral = indir_integral * 2 * np.pi
        diffuse_albedo, roughness = mat['sg_diffuse_albedo'], mat['sg_roughness']
        zero3 = torch.zeros_like(normals)
        normal_map = torch.where(valid[:, None], mat['sg_normal_map'].detach(), zero3)
        emb = self.shadow_embed(points.detach())
        with torch.set_grad_enabled(self.is_training and torch.is_grad_enabled()):
            diffuse_vis = shadow_logits(self.shadow_net, emb, M, n_active=n_act)
            normal_new = wn_mlp(self.normal_net, emb, n_active=n_act)
        normal_new = normal_new / torch.clamp(normal_new.norm(dim=-1, keepdim=True), 1e-4)
        normal_new = torch.where(valid[:, None], normal_new, zero3)
        diffuse_vis = torch.softmax(diffuse_vis, -1)[..., 1]
        prefit = self.prefit_option()
        sg = sg_render.render_with_all_sg(points=points.detach(),
                                          normal=normal_new if self.cur_iter > 1000 else normal_map,
                                          viewdirs=view_dirs, lgtSGs=lgtSGs, indir_integral=indir_integral,
                                          specular_reflectance=mat['sg_specular_reflectance'].abs(),
                                          roughness=roughness, diffuse_albedo=diffuse_albedo,
                                          indir_lgtSGs=indir_lgtSGs, VisModel=model.visibility_network, fun_spec=False,
                                          lin_diff=True, testing=not self.is_training, metallic=None,
                                          diffuse_vis=diffuse_vis, prefit=prefit, argmax_vis=False, valid=valid,
                                          diffuse_presampled=diffuse_presampled)
        sg["sg_rgb"] = sg["sg_diffuse_rgb"] * diffuse_albedo / np.pi + sg["sg_specular_rgb"]
        sg["indir_rgb"] = sg["indir_diffuse_rgb"] * diffuse_albedo / np.pi + sg["indir_specular_rgb"]
        supervise = sg['supervise']
        if self.white_light and prefit != "warmup":
            supervise = supervise + white_loss(lgtSGs)
        count = valid.sum().clamp(min=1) * 3
        supervise = supervise + ((normal_map - normal_new) ** 2).sum() / count          # both are 0 on the other rows
        ret.update(sg)
        ret.update({'diffuse_albedo': diffuse_albedo, 'roughness': roughness, 'metallic': mat['sg_metallic'],
                    'normal_map': normal_new, 'gradient_error': supervise,
                    'random_xi_roughness': mat['random_xi_roughness'],
                    'random_xi_metallic': mat['random_xi_metallic'],
                    'random_xi_diffuse_albedo': mat['random_xi_diffuse_albedo']})
        return ret

    # ---- the step loss (train_cesr.py:387-430); loss_fn is an InvLoss
    def pbr_step(self, loss_fn, model_outputs, ground_truth):
        loss = 0.
        out = {}
        if self.cur_iter > 500:
            out = loss_fn(model_outputs, ground_truth, mat_model=self.model.envmap_material_network, train_idr=False,
                          train_spec=self.train_spec, hdr_fn=self.model.gamma.hdr_shift.hdr2ldr)
            smooth_w, kl_w = self.weights["project" if self.prefit_option() == "project" else "explore"]
            loss = out["loss"] + out["kl_loss"] * kl_w + out["latent_smooth_loss"] * smooth_w
        return loss + model_outputs["gradient_error"], out


def shadow_logits(net, emb, M=128, n_active=None):
    """shadow_net on every (point, lobe) pair: rows [PE10(x_i) | onehot(m)], i-major (train_cesr.py:492-501).
    n_active: device int32 [1] count of leading POINTS that carry data (fixed-capacity mode)."""
    n = emb.shape[0]
    x = torch.cat([emb[:, None, :].expand(-1, M, -1), torch.eye(M, device=emb.device)[None].expand(n, -1, -1)], -1)
    return wn_mlp(net, x.reshape(n * M, -1), n_active=None if n_active is None else n_active * M)
