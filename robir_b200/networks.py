"""Parameter containers with the reference's module tree / state_dict keys (134 tensors), so
that reference checkpoints load here and vice versa.  The heavy evaluation paths go through the CUDA operators in
``ops``; the small material / indirect networks (<2 % of the step's FLOPs) are plain library GEMMs.

Reference classes mirrored: SDFNetwork / RenderingNetwork / SingleVarianceNetwork / NeuSModel / ImplicitNetworkMy
(model/neus_model.py:312-438,489-560,644-650,682-884), SparseAE / EnvmapMaterialNetwork
(model/sg_envmap_material.py:40-275), IndirctIllumNetwork / VisNetwork (model/implicit_differentiable_renderer.py:
170-258), GammaCorrect / ACESToneMapping (model/color_correction.py:7-137).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, rng

FUSED_MLP = True   # CUDA tensors take the fused-kernel path (csrc/mlp.cu); False = plain library GEMMs (debug)


def positional_encoding(x, n_freq):
    out = [x]
    for l in range(n_freq):
        f = float(2 ** l)
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, -1)


def integrated_positional_encoding(x, n_deg=10, var=1e-5):
    scales = torch.exp2(torch.arange(n_deg, device=x.device, dtype=x.dtype))   # no H2D copy: CUDA-graph capturable
    y = (x[:, None, :] * scales[:, None]).reshape(x.shape[0], -1)            # [n, 3*n_deg], degree-major
    y_var = (var * scales ** 2)[:, None].expand(n_deg, x.shape[1]).reshape(1, -1)
    yy = torch.cat([y, y + 0.5 * math.pi], -1)
    vv = torch.cat([y_var, y_var], -1)
    lim = 100 * math.pi
    safe = torch.where(yy.abs() < lim, yy, yy % lim)
    return torch.exp(-0.5 * vv) * torch.sin(safe)


class _WNLinear(nn.Module):
    """weight_norm(dim=0)-parametrised linear layer holding weight_g / weight_v / bias like torch's legacy hook."""

    def __init__(self, in_f, out_f):
        super().__init__()
        w = torch.empty(out_f, in_f).uniform_(-1, 1) / math.sqrt(in_f)
        self.weight_g = nn.Parameter(w.norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(w)
        self.bias = nn.Parameter(torch.zeros(out_f))

    def folded(self):
        return self.weight_g * self.weight_v / self.weight_v.norm(dim=1, keepdim=True)


class SDFNetwork(nn.Module):
    def __init__(self):
        super().__init__()
        dims = [63] + [256] * 8 + [257]
        for l in range(9):
            setattr(self, "lin%d" % l, _WNLinear(dims[l], dims[l + 1] - 63 if l + 1 == 4 else dims[l + 1]))


class RenderingNetwork(nn.Module):
    def __init__(self):
        super().__init__()
        dims = [289, 256, 256, 256, 256, 3]
        for l in range(5):
            setattr(self, "lin%d" % l, _WNLinear(dims[l], dims[l + 1]))

    def forward(self, points, normals, view_dirs, feats):
        x = torch.cat([points, positional_encoding(view_dirs, 4), normals, feats], -1)
        for l in range(5):
            lin = getattr(self, "lin%d" % l)
            x = F.linear(x, lin.folded(), lin.bias)
            if l < 4:
                x = torch.relu(x)
        return torch.sigmoid(x)


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val=0.3):
        super().__init__()
        self.variance = nn.Parameter(torch.tensor(init_val))


class NeuSModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.color_network = RenderingNetwork()
        self.sdf_network = SDFNetwork()
        self.deviation_network = SingleVarianceNetwork(0.3)


class ImplicitNetworkMy(nn.Module):
    """f(p) = SDFNetwork(2p) / 2; value, gradient and features come from the fused CUDA kernel (ops.sdf_eval)."""

    def __init__(self):
        super().__init__()
        self.neus_model = NeuSModel()
        self._w = ops.SdfWeights(self.neus_model.sdf_network)

    def forward(self, points):
        """[k,3] -> [k,257] (sdf, features), all halved like the reference (neus_model.py:788-792)."""
        if points.numel() == 0:
            return torch.ones_like(points)
        sdf, _, feat = ops.sdf_eval(self._w, points, want_feat=True)
        return torch.cat([sdf[:, None], feat], -1)

    def sdf(self, points):
        """[k,3] -> [k]: what every tracer calls (``lambda x: implicit_network(x)[:, 0]``) without the feature GEMM."""
        return ops.sdf_eval(self._w, points)[0]

    def gradient(self, points):
        """[k,3] -> [k,1,3]; graph-free (hard part 6)."""
        if points.numel() == 0:
            return torch.ones_like(points)
        return ops.sdf_eval(self._w, points, want_grad=True)[1].unsqueeze(1)

    def sdf_and_normal(self, points):
        sdf, grad, _ = ops.sdf_eval(self._w, points, want_grad=True)
        return sdf, grad

    # ---- NeuS-coordinate evaluation used by the secondary-ray radiance (neus_model.py:745-752, 828-884) ----------------
    def neus_forward(self, pts_neus, dirs, return_grad=False):
        """NeuSModel.forward: colour(x, grad sdf(x), dirs, feat) and sdf at NeuS-coordinate points (no autograd)."""
        sdf, grad, feat = ops.sdf_eval(self._w, pts_neus, in_scale=1.0, sdf_scale=1.0, feat_scale=1.0, want_grad=True,
                                       want_feat=True)
        if self.__dict__.get("_color_chain") is None:
            net = self.neus_model.color_network
            self.__dict__["_color_chain"] = ops.MlpChain.from_weightnorm([getattr(net, "lin%d" % l) for l in range(5)],
                                                                          "relu", "raw")
        x = torch.cat([pts_neus, positional_encoding(dirs, 4), grad, feat], -1)
        color = torch.sigmoid(ops.fused_mlp(self.__dict__["_color_chain"], x))
        if return_grad:
            return color, sdf[:, None], grad
        return color, sdf[:, None]

    def borrow_color(self, points, view_dirs):
        """16-sample NeuS micro volume render around a surface point (neus_model.py:828-869)."""
        vd = -view_dirs / torch.norm(view_dirs, dim=-1, keepdim=True)
        n_samp = 16
        t = torch.linspace(-0.01, 0.05, n_samp, device=points.device)[:, None]
        p = points[:, None, :] * 2 + vd[:, None, :] * t
        d = vd[:, None, :].expand(-1, n_samp, -1)
        color, sdf = self.neus_forward(p.reshape(-1, 3), d.reshape(-1, 3))
        color, sdf = color.view(-1, n_samp, 3), sdf.view(-1, n_samp, 1)
        inv_s = torch.exp(self.neus_model.deviation_network.variance * 10.0).clip(1e-6, 1e6)
        nxt = torch.cat([sdf[:, 1:], sdf[:, -1:]], 1)
        prv = torch.cat([sdf[:, :-1], sdf[:, -1:]], 1)
        prev_cdf, next_cdf = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
        alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).reshape(-1, n_samp).clip(0.0, 1.0)
        trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
        return (color * (alpha * trans)[:, :, None]).sum(1)

    def batch_borrow_color(self, points, view_dirs, batch=8192):
        if points.shape[0] == 0:
            return torch.zeros_like(points)
        with torch.no_grad():
            return torch.cat([self.borrow_color(points[i:i + batch], view_dirs[i:i + batch])
                              for i in range(0, points.shape[0], batch)], 0)


def _mlp(dims, act):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(act)
    return nn.Sequential(*layers)


class SparseAE(nn.Module):
    """Encoder in->512x4->32, latent activation, decoder 32->128x2->out; noisy twin output (random_xi)."""

    def __init__(self, in_dim, out_dim, smooth_on_latent=True, out_act=torch.sigmoid, latent_dim=32, high_lr=False):
        super().__init__()
        self.actv_fn = nn.LeakyReLU(0.2)
        self.brdf_encoder_layer = _mlp([in_dim, 512, 512, 512, 512, latent_dim], self.actv_fn)
        self.brdf_decoder_layer = _mlp([latent_dim, 128, 128, out_dim], self.actv_fn)
        self.smooth_on_latent = smooth_on_latent
        self.out_act = out_act
        self.lc_act = torch.sigmoid
        self.latent_dim = latent_dim
        self.var = None  # CESR latent dropout (train_cesr.py:639-641); zeros in the PBR stage

    def encode(self, x):
        z = self.brdf_encoder_layer(x)
        if self.var is not None:
            z = z * (1 - self.var.to(z.device))
        return z

    # ---- fused CUDA path: the encoding + every Linear/LeakyReLU chain is one kernel launch (csrc/mlp.cu) -------------
    def _chain(self, which, in_mode):
        key = "_robir_%s_%s" % (which, in_mode)
        ch = self.__dict__.get(key)
        if ch is None:
            seq = self.brdf_encoder_layer if which == "enc" else self.brdf_decoder_layer
            ch = ops.MlpChain.from_sequential(seq, "leaky", in_mode)
            self.__dict__[key] = ch
        return ch

    def _wants_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def encode_points(self, points, in_mode="pe10", extra=None, noise=None, train=None, segments=1):
        """encoder(embed(points) (+ 0.02 noise)) -> latent pre-activation [n, 32]"""
        train = self._wants_grad() if train is None else train
        z = ops.fused_mlp(self._chain("enc", in_mode), points, extra=extra, noise=noise, noise_scale=0.02,
                          want_param_grad=train, segments=segments)
        if self.var is not None:
            z = z * (1 - self.var.to(z.device))
        return z

    def forward_points(self, points, in_mode="pe10", extra=None, train=None, noisy_only=False):
        """SparseAE.forward on raw points (model/sg_envmap_material.py:74-94) -> (y, y_noisy, z_clean)."""
        train = self._wants_grad() if train is None else train
        n = points.shape[0]
        dec = self._chain("dec", "raw")
        if self.smooth_on_latent:
            z = self.encode_points(points, in_mode, extra, None, train)
            lc = self.lc_act(z)
            lc_r = lc + rng.randn(lc.shape, lc.device) * 0.01
            y2 = ops.fused_mlp(dec, torch.cat([lc, lc_r], 0), want_param_grad=train, segments=2)
            y, y_r = y2[:n], y2[n:]
        else:
            in_dim = self.brdf_encoder_layer[0].in_features
            noise = rng.randn((n, in_dim), points.device)
            if noisy_only:
                z = None
                lc_r = self.lc_act(self.encode_points(points, in_mode, extra, noise, train))
                y = None
                y_r = ops.fused_mlp(dec, lc_r, want_param_grad=train)
            else:
                pts2 = torch.cat([points, points], 0)
                ex2 = torch.cat([extra, extra], 0) if extra is not None else None
                z2 = self.encode_points(pts2, in_mode, ex2, torch.cat([torch.zeros_like(noise), noise], 0), train,
                                        segments=2)
                z = z2[:n]
                y2 = ops.fused_mlp(dec, self.lc_act(z2), want_param_grad=train, segments=2)
                y, y_r = y2[:n], y2[n:]
        if self.out_act is not None:
            y = self.out_act(y) if y is not None else None
            y_r = self.out_act(y_r)
        return y, y_r, z

    def brdf_points(self, points, train=None):
        """The BRDF auto-encoder of EnvmapMaterialNetwork on raw points (sg_envmap_material.py:214-232): encoder chain,
        fused latent pair, decoder chain on the doubled batch, fused output head -> the six material tensors + z."""
        train = self._wants_grad() if train is None else train
        sig = (torch.sigmoid, F.sigmoid)
        if not self.smooth_on_latent or self.latent_dim != 32 or self.lc_act not in sig or self.out_act not in sig:
            y, y_r, z = self.forward_points(points, "pe10", train=train)
            return dict(z=z, sg_roughness=y[..., 3:4] * 0.9 + 0.09, sg_metallic=y[..., 4:5] * 0.99 + 0.01,
                        sg_diffuse_albedo=y[..., :3], random_xi_roughness=y_r[..., 3:4] * 0.9 + 0.09,
                        random_xi_diffuse_albedo=y_r[..., :3], random_xi_metallic=y_r[..., 4:5])
        z = self.encode_points(points, "pe10", None, None, train)
        lc2 = ops.latent_pair(z, rng.randn(z.shape, z.device))
        y2 = ops.fused_mlp(self._chain("dec", "raw"), lc2, want_param_grad=train, segments=2)
        alb, rough, metal, alb_r, rough_r, metal_r = ops.brdf_head(y2)
        return dict(z=z, sg_roughness=rough, sg_metallic=metal, sg_diffuse_albedo=alb, random_xi_roughness=rough_r,
                    random_xi_diffuse_albedo=alb_r, random_xi_metallic=metal_r)

    def forward(self, x):
        lc = self.lc_act(self.encode(x))
        y = self.brdf_decoder_layer(lc)
        if self.smooth_on_latent:
            lc_r = lc + rng.randn(lc.shape, lc.device) * 0.01
        else:
            lc_r = self.lc_act(self.encode(x + rng.randn(x.shape, x.device) * 0.02))
        y_r = self.brdf_decoder_layer(lc_r)
        if self.out_act is not None:
            y, y_r = self.out_act(y), self.out_act(y_r)
        return y, y_r


class EnvmapMaterialNetwork(nn.Module):
    def __init__(self, multires=10, brdf_encoder_dims=None, brdf_decoder_dims=None, num_lgt_sgs=128, upper_hemi=False,
                 specular_albedo=0.05, latent_dim=32):
        super().__init__()
        assert multires == 10, "the shipped configs use multires = 10 (confs_sg/*.conf)"
        self.numLgtSGs = num_lgt_sgs
        self.envmap = None
        self.upper_hemi = upper_hemi
        self.brdf_encoder_layer = SparseAE(63, 5, out_act=None)
        self.spec_brdf_encoder_layer = SparseAE(63, 5, high_lr=True)
        self.normal_decoder_layer = SparseAE(60, 3, out_act=None, smooth_on_latent=False)
        self.specular_reflectance = nn.Parameter(torch.full((1, 1), float(specular_albedo)))
        from .synthetic import synthetic_light_sgs
        self.lgtSGs = nn.Parameter(synthetic_light_sgs(torch.Generator().manual_seed(0), num_lgt_sgs))

    def forward(self, points, train_spec=False, train_norm=False):
        fused = points.is_cuda and FUSED_MLP
        ret = {}
        if not fused:
            pts_ipe = integrated_positional_encoding(points, 10, 1e-5)
            emb = positional_encoding(points, 10)
        nm = None
        if not train_norm:
            if fused:
                # the BRDF and the normal auto-encoders only share the input points: two parallel branches (host call
                # order = random-draw order of the reference: BRDF latent noise, then normal input noise)
                head, (nm, nm_r, _) = ops.fork_join([
                    lambda: self.spec_brdf_encoder_layer.brdf_points(points.detach(), train=None if train_spec else False),
                    lambda: self.normal_decoder_layer.forward_points(points.detach(), "ipe10")], tag="mat")
                self._last_spec_latent = head.pop("z")   # reused by the KL term of the loss (same points, same encoder)
                if train_spec is False:
                    head = {k: v.detach() for k, v in head.items()}
                ret.update(head)
            else:
                brdf, brdf_r = self.spec_brdf_encoder_layer(emb)
                if train_spec is False:
                    brdf, brdf_r = brdf.detach(), brdf_r.detach()
                ret.update(sg_roughness=brdf[..., 3:4] * 0.9 + 0.09, sg_metallic=brdf[..., 4:5] * 0.99 + 0.01,
                           sg_diffuse_albedo=brdf[..., :3], random_xi_roughness=brdf_r[..., 3:4] * 0.9 + 0.09,
                           random_xi_diffuse_albedo=brdf_r[..., :3], random_xi_metallic=brdf_r[..., 4:5])
        if fused:
            if nm is None:
                nm, nm_r, _ = self.normal_decoder_layer.forward_points(points.detach(), "ipe10")
        else:
            nm, nm_r = self.normal_decoder_layer(pts_ipe)
        ret["sg_normal_map"] = nm / torch.clamp(nm.norm(dim=-1, keepdim=True), 1e-4)
        ret["random_xi_normal"] = nm_r / torch.clamp(nm_r.norm(dim=-1, keepdim=True), 1e-4)
        if train_norm:
            return dict(sg_normal_map=ret["sg_normal_map"], random_xi_normal=ret["random_xi_normal"])
        lgt = self.lgtSGs
        if self.upper_hemi:
            lgt = torch.cat((lgt[..., :1], torch.abs(lgt[..., 1:2]), lgt[..., 2:]), dim=-1)
        ret["sg_lgtSGs"] = lgt
        ret["sg_specular_reflectance"] = self.specular_reflectance
        return ret


class IndirctIllumNetwork(nn.Module):
    def __init__(self, multires=10, dims=(512, 512, 512, 512), num_lgt_sgs=24, no_hdr=False):
        super().__init__()
        assert multires == 10 and not no_hdr
        self.num_lgt_sgs = num_lgt_sgs
        self.lobe_layer = _mlp([64] + list(dims) + [num_lgt_sgs * 6], nn.ReLU())
        self.integral_layer = SparseAE(64, 3, out_act=None, smooth_on_latent=False)
        self.integral_layer.lc_act = F.softplus
        self._lobe_chain = None
        # PBR stage: these weights are not optimised (train_pbr.py:104-105) -> only the input gradient (hdr_shift)
        # is propagated; the Vis stage sets train_weights = True
        self.train_weights = False

    def forward(self, points, hdr_shift):
        if points.is_cuda and FUSED_MLP:
            if self._lobe_chain is None:
                self._lobe_chain = ops.MlpChain.from_sequential(self.lobe_layer, "relu", "pe10_extra")
            train = self.train_weights and torch.is_grad_enabled()
            pts = points.detach()
            # lobe network and integral auto-encoder are independent: two parallel branches (forward and backward)
            out, (_, env_r, _) = ops.fork_join([
                lambda: ops.fused_mlp(self._lobe_chain, pts, extra=hdr_shift, want_param_grad=train),
                lambda: self.integral_layer.forward_points(pts, "pe10_extra", extra=hdr_shift, train=train,
                                                           noisy_only=True)], tag="indir")
            out = out.reshape(pts.shape[0], self.num_lgt_sgs, 6)
            return ops.decode_lobes(out), torch.abs(env_r)
        x = torch.cat([positional_encoding(points, 10), hdr_shift], -1)
        out = self.lobe_layer(x).reshape(x.shape[0], self.num_lgt_sgs, 6)
        env_int = torch.abs(self.integral_layer(x)[1])
        return self._decode_lobes(out), env_int

    @staticmethod
    def _decode_lobes(out):
        ang = torch.sigmoid(out[..., :2])
        theta, phi = ang[..., :1] * 2 * np.pi, ang[..., 1:2] * np.pi
        lobes = torch.cat([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], -1)
        return torch.cat([lobes, torch.sigmoid(out[..., 2:3]) * 30 + 0.1, torch.relu(out[..., 3:])], -1)


class VisNetwork(nn.Module):
    def __init__(self, points_multires=10, dirs_multires=10, dims=(256, 256, 256, 256)):
        super().__init__()
        assert points_multires == 10 and dirs_multires == 10 and tuple(dims) == (256,) * 4
        self.vis_layer = _mlp([126] + list(dims) + [2], nn.ReLU())

    def forward(self, points, view_dirs):
        """Plain logits [k,2] (used where the reference calls the network directly, e.g. trace_radiance :632-634)."""
        if points.is_cuda and FUSED_MLP:
            if self.__dict__.get("_chain") is None:
                self.__dict__["_chain"] = ops.MlpChain.from_sequential(self.vis_layer, "relu", "pe10x2")
            train = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
            return ops.fused_mlp(self.__dict__["_chain"], torch.cat([points, view_dirs], -1).detach(),
                                 want_param_grad=train)
        return self.vis_layer(torch.cat([positional_encoding(points, 10), positional_encoding(view_dirs, 10)], -1))


def aces(x):
    return x * (2.51 * x + 0.03) / (x * (2.43 * x + 0.59) + 0.14)


def aces_inverse(x):
    return ((0.59 * x - 0.03) + torch.sqrt((0.59 * x - 0.03) ** 2 + 4 * (2.51 - 2.43 * x) * 0.14 * x)) / (
        2 * (2.51 - 2.43 * x))


class ACESToneMapping(nn.Module):
    def __init__(self, hdr_mode=0):
        super().__init__()
        assert hdr_mode == 0, "shipped configs use hdr_mode = 0 (confs_sg/*.conf:67)"
        self.adapt_illum = nn.Parameter(torch.tensor(0.0))

    def as_input(self):
        return torch.clamp(self.adapt_illum * 10 + 0.5, 0, 1).view(1, 1)

    def make_shift(self, shift):
        if shift is None:
            shift = self.as_input()
        if not isinstance(shift, torch.Tensor):
            shift = torch.tensor(shift, device=self.adapt_illum.device)
        if shift.dim() == 0:
            shift = shift[None]
        return torch.clamp(shift, 1e-4, 1)

    def hdr2ldr(self, x, raw_shift=None):
        return aces(x) / self.make_shift(raw_shift) ** 0.2

    def ldr2hdr(self, x, raw_shift=None):
        return aces_inverse(x * self.make_shift(raw_shift) ** 0.2)


class GammaCorrect(nn.Module):
    def __init__(self, gamma=1.0, hdr_mode=0):
        super().__init__()
        self.gamma = nn.Parameter(torch.tensor(float(gamma)))
        self.indir_coef = nn.Parameter(torch.tensor(1.0))
        self.dir_coef = nn.Parameter(torch.tensor(2.0))
        self.coef = nn.Parameter(torch.tensor(1.0))
        self.hdr_shift = ACESToneMapping(hdr_mode)

    def forward(self, x):
        return torch.pow(x, 1 / self.gamma)
