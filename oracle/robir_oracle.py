"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch, fp32) restatement of RobIR's per-ray rendering hot path, written as
pure functions over a flat ``state_dict`` (the 134 tensors of the reference ``IDRNetwork.state_dict()``) with every
random draw passed in explicitly, so that reference, oracle and CUDA product can be compared on
identical inputs.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package ``robir_b200`` never does.

Parity status: PINNED against the unmodified reference run through ``oracle/ref_shim.py`` in the build container
(``tests/test_oracle_vs_reference.py``, skipped where /root/reference is absent) and against the committed golden
vectors in ``tests/golden/`` that ``tests/golden/make_golden.py`` produced from the reference.  The reference ships
no golden vectors or tests of its own.

Each function cites the reference file:line it follows (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

TINY = 1e-6  # model/sg_render.py:6


# ----------------------------------------------------------------------------------------------------------------------
# encodings  (model/embedder.py:12-38, model/neus_model.py:14-94,136-181)
# ----------------------------------------------------------------------------------------------------------------------
def pe(x, n_freq):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] -- model/embedder.py:24-38."""
    out = [x]
    freqs = 2.0 ** torch.linspace(0.0, n_freq - 1, n_freq)
    for f in freqs:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def ipe(x, n_deg=10, var=1e-5):
    """Integrated PE with isotropic covariance -- model/neus_model.py:25-94 via model/embedder.py:58-61.
    Output: [exp(-y_var/2) sin(y) (3*n_deg) | exp(-y_var/2) sin(y + pi/2) (3*n_deg)], no raw x."""
    d = x.shape[-1]
    basis = torch.cat([2 ** i * torch.eye(d) for i in range(n_deg)], 1)  # [d, d*n_deg]
    y = x @ basis
    cov = (torch.eye(d) * var)[None].expand(x.shape[0], -1, -1)
    y_var = torch.sum((cov @ basis) * basis, -2)
    yy = torch.cat([y, y + 0.5 * torch.pi], -1)
    vv = torch.cat([y_var, y_var], -1)
    t = 100 * torch.pi
    safe = torch.where(torch.abs(yy) < t, yy, yy % t)  # neus_model.py:15-16
    return torch.exp(-0.5 * vv) * torch.sin(safe)


# ----------------------------------------------------------------------------------------------------------------------
# NeuS SDF / colour networks (model/neus_model.py:312-438, 489-560, 755-884)
# ----------------------------------------------------------------------------------------------------------------------
def wn_weight(sd, key):
    """nn.utils.weight_norm(dim=0): W[o,:] = g[o] * v[o,:] / ||v[o,:]||."""
    v = sd[key + ".weight_v"]
    g = sd[key + ".weight_g"]
    return g * v / v.norm(dim=1, keepdim=True)


SDF_PREFIX = "implicit_network.neus_model.sdf_network"
COLOR_PREFIX = "implicit_network.neus_model.color_network"


def softplus100(x):
    return F.softplus(x, beta=100)


def sdf_network(sd, pts, chunk=1024):
    """SDFNetwork.forward (neus_model.py:385-417): PE(10) -> 9 weight-normed linears, skip at 4, softplus(100).
    pts are in NeuS coordinates.  Returns [n, 257].  The reference evaluates in 1024-row chunks; chunking changes
    nothing numerically on CPU beyond GEMM blocking, we keep it for faithfulness."""
    if pts.numel() == 0:
        return torch.ones_like(pts)
    emb = pe(pts, 10)
    res = []
    for c in range(emb.shape[0] // chunk + 1):
        e = emb[chunk * c: chunk * (c + 1)]
        x = e
        for l in range(9):
            if l == 4:
                x = torch.cat([x, e], 1) / np.sqrt(2)
            x = F.linear(x, wn_weight(sd, "%s.lin%d" % (SDF_PREFIX, l)), sd["%s.lin%d.bias" % (SDF_PREFIX, l)])
            if l < 8:
                x = softplus100(x)
        res.append(x)
    return torch.cat(res, 0)


def wn_mlp(sd, prefix, x, n_lin=9, skip_in=(4,), chunk=1024, prefix_dot=True):
    """SDFNetwork.forward with multires=0 (neus_model.py:385-417), the form of the CESR stage's shadow_net / normal_net
    (training/train_cesr.py:106-110): ``n_lin`` weight-normed linears under ``prefix.lin{l}``, the layer input is
    cat([x, inputs]) / sqrt(2) at the skip layers, softplus(100) between layers, evaluated in 1024-row chunks."""
    if x.numel() == 0:
        return torch.ones_like(x)
    shape = list(x.shape[:-1]) + [-1]
    inputs = x.reshape(-1, x.shape[-1])
    res = []
    for c in range(inputs.shape[0] // chunk + 1):
        e = inputs[chunk * c: chunk * (c + 1)]
        h = e
        for l in range(n_lin):
            if l in skip_in:
                h = torch.cat([h, e], 1) / np.sqrt(2)
            key = ("%s.lin%d" if prefix_dot else "%slin%d") % (prefix, l)
            h = F.linear(h, wn_weight(sd, key), sd[key + ".bias"])
            if l < n_lin - 1:
                h = softplus100(h)
        res.append(h)
    return torch.cat(res, 0).reshape(shape)


def implicit_forward(sd, pts):
    """ImplicitNetworkMy.forward (neus_model.py:785-792): net(2 p) / 2 on all 257 channels."""
    return sdf_network(sd, pts * 2.0) / 2.0


def implicit_gradient(sd, pts):
    """ImplicitNetworkMy.gradient (neus_model.py:803-818) -> [n, 1, 3] (graph-free here, SURVEY hard-part 6)."""
    if pts.numel() == 0:
        return torch.ones_like(pts)
    with torch.enable_grad():
        x = pts.detach().clone().requires_grad_(True)
        y = implicit_forward(sd, x)[:, :1]
        g = torch.autograd.grad(y, x, torch.ones_like(y))[0]
    return g.detach().unsqueeze(1)


def color_network(sd, pts_neus, normals, view_dirs, feats):
    """RenderingNetwork.forward, mode 'idr' (neus_model.py:536-560): cat(points, PE4(view), normals, feat) -> sigmoid."""
    x = torch.cat([pts_neus, pe(view_dirs, 4), normals, feats], -1)
    for l in range(5):
        x = F.linear(x, wn_weight(sd, "%s.lin%d" % (COLOR_PREFIX, l)), sd["%s.lin%d.bias" % (COLOR_PREFIX, l)])
        if l < 4:
            x = torch.relu(x)
    return torch.sigmoid(x)


def neus_forward(sd, pts_neus, dirs):
    """NeuSModel.forward (neus_model.py:745-752): colour(x, grad sdf(x), dirs, feat) and sdf, NeuS coordinates."""
    out = sdf_network(sd, pts_neus)
    sdf, feat = out[:, :1], out[:, 1:]
    with torch.enable_grad():
        x = pts_neus.detach().clone().requires_grad_(True)
        y = sdf_network(sd, x)[:, :1]
        g = torch.autograd.grad(y, x, torch.ones_like(y))[0].detach()
    return color_network(sd, pts_neus, g, dirs, feat), sdf


def borrow_color(sd, points, view_dirs):
    """ImplicitNetworkMy.borrow_color + volume_render (neus_model.py:828-869): 16-sample NeuS micro volume render."""
    vd = -view_dirs / torch.norm(view_dirs, dim=-1, keepdim=True)
    n_samp = 16
    t = torch.linspace(-0.01, 0.05, n_samp)[:, None]
    p = points[:, None, :] * 2 + vd[:, None, :] * t
    d = vd[:, None, :].expand(-1, n_samp, -1)
    color, sdf = neus_forward(sd, p.reshape(-1, 3), d.reshape(-1, 3))
    color = color.view(-1, n_samp, 3)
    sdf = sdf.view(-1, n_samp, 1)
    inv_s = torch.exp(sd["implicit_network.neus_model.deviation_network.variance"] * 10.0).clip(1e-6, 1e6)
    nxt = torch.cat([sdf[:, 1:], sdf[:, -1:]], 1)
    prv = torch.cat([sdf[:, :-1], sdf[:, -1:]], 1)
    prev_cdf = torch.sigmoid(prv * inv_s)
    next_cdf = torch.sigmoid(nxt * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).reshape(-1, n_samp).clip(0.0, 1.0)
    trans = torch.cumprod(torch.cat([torch.ones(alpha.shape[0], 1), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return (color * (alpha * trans)[:, :, None]).sum(1)


def batch_borrow_color(sd, points, view_dirs, batch=8192):
    """neus_model.py:871-884."""
    if points.shape[0] == 0:
        return torch.zeros_like(points)
    with torch.no_grad():
        return torch.cat([borrow_color(sd, points[i:i + batch], view_dirs[i:i + batch])
                          for i in range(0, points.shape[0], batch)], 0)


# ----------------------------------------------------------------------------------------------------------------------
# material / light / visibility networks
# ----------------------------------------------------------------------------------------------------------------------
def _seq(sd, prefix, idxs, x, act, last_act=False):
    for n, i in enumerate(idxs):
        x = F.linear(x, sd["%s.%d.weight" % (prefix, i)], sd["%s.%d.bias" % (prefix, i)])
        if n < len(idxs) - 1 or last_act:
            x = act(x)
    return x


def sparse_ae_encode(sd, prefix, x):
    """SparseAE.encode (sg_envmap_material.py:96-99), var == 0."""
    return _seq(sd, prefix + ".brdf_encoder_layer", [0, 2, 4, 6, 8], x, lambda t: F.leaky_relu(t, 0.2))


def sparse_ae_decode(sd, prefix, z):
    return _seq(sd, prefix + ".brdf_decoder_layer", [0, 2, 4], z, lambda t: F.leaky_relu(t, 0.2))


def sparse_ae(sd, prefix, x, noise, smooth_on_latent=True, out_act=torch.sigmoid, lc_act=torch.sigmoid):
    """SparseAE.forward (sg_envmap_material.py:74-94).  noise: randn of the latent shape [n,32] if smooth_on_latent
    else of the (already embedded) input shape."""
    lc = lc_act(sparse_ae_encode(sd, prefix, x))
    y = sparse_ae_decode(sd, prefix, lc)
    if smooth_on_latent:
        lc_r = lc + noise * 0.01
    else:
        lc_r = lc_act(sparse_ae_encode(sd, prefix, x + noise * 0.02))
    y_r = sparse_ae_decode(sd, prefix, lc_r)
    if out_act is not None:
        y, y_r = out_act(y), out_act(y_r)
    return y, y_r


MAT = "envmap_material_network"


def envmap_material(sd, points, noise_brdf_latent, noise_normal_in, train_norm=False):
    """EnvmapMaterialNetwork.forward with train_spec=True (sg_envmap_material.py:188-247)."""
    pts_ipe = ipe(points, 10, 1e-5)
    emb = pe(points, 10)
    ret = {}
    if not train_norm:
        brdf, brdf_r = sparse_ae(sd, MAT + ".spec_brdf_encoder_layer", emb, noise_brdf_latent)
        ret.update(sg_roughness=brdf[..., 3:4] * 0.9 + 0.09, sg_metallic=brdf[..., 4:5] * 0.99 + 0.01,
                   sg_diffuse_albedo=brdf[..., :3], random_xi_roughness=brdf_r[..., 3:4] * 0.9 + 0.09,
                   random_xi_diffuse_albedo=brdf_r[..., :3], random_xi_metallic=brdf_r[..., 4:5])
    nm, nm_r = sparse_ae(sd, MAT + ".normal_decoder_layer", pts_ipe, noise_normal_in, smooth_on_latent=False,
                         out_act=None)
    ret["sg_normal_map"] = nm / torch.clamp(nm.norm(dim=-1, keepdim=True), 1e-4)
    ret["random_xi_normal"] = nm_r / torch.clamp(nm_r.norm(dim=-1, keepdim=True), 1e-4)
    ret["sg_lgtSGs"] = sd[MAT + ".lgtSGs"]
    ret["sg_specular_reflectance"] = sd[MAT + ".specular_reflectance"]
    return ret


IND = "indirect_illum_network"


def indirect_illum(sd, points, hdr_shift, noise_in):
    """IndirctIllumNetwork.forward (implicit_differentiable_renderer.py:199-222). noise_in: randn [n,64]."""
    x = torch.cat([pe(points, 10), hdr_shift], -1)
    n = x.shape[0]
    out = _seq(sd, IND + ".lobe_layer", [0, 2, 4, 6, 8], x, torch.relu).reshape(n, -1, 6)
    ang = torch.sigmoid(out[..., :2])
    theta, phi = ang[..., :1] * 2 * np.pi, ang[..., 1:2] * np.pi
    lobes = torch.cat([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], -1)
    lam = torch.sigmoid(out[..., 2:3]) * 30 + 0.1
    mu = torch.relu(out[..., 3:])
    sgs = torch.cat([lobes, lam, mu], -1)
    _, env_r = sparse_ae(sd, IND + ".integral_layer", x, noise_in, smooth_on_latent=False, out_act=None,
                         lc_act=F.softplus)
    return sgs, torch.abs(env_r)


VIS = "visibility_network.vis_layer"


def vis_network(sd, points, dirs):
    """VisNetwork.forward (implicit_differentiable_renderer.py:250-258): logits [k,2]."""
    return _seq(sd, VIS, [0, 2, 4, 6, 8], torch.cat([pe(points, 10), pe(dirs, 10)], -1), torch.relu)


# ----------------------------------------------------------------------------------------------------------------------
# SG math (model/sg_render.py:62-108)
# ----------------------------------------------------------------------------------------------------------------------
def norm_axis(x):
    return x / (torch.norm(x, dim=-1, keepdim=True) + TINY)


def hemisphere_int(lam, cos_beta):
    lam = lam + TINY
    inv = 1.0 / lam
    t = torch.sqrt(lam) * (1.6988 + 10.8438 * inv) / (1.0 + 6.2201 * inv + 10.2415 * inv * inv)
    inv_a = torch.exp(-t)
    m = (cos_beta >= 0).float()
    inv_b = torch.exp(-t * torch.clamp(cos_beta, min=0.0))
    s1 = (1.0 - inv_a * inv_b) / (1.0 - inv_a + inv_b - inv_a * inv_b)
    b = torch.exp(t * torch.clamp(cos_beta, max=0.0))
    s2 = (b - inv_a) / ((1.0 - inv_a) * (b + 1.0))
    s = m * s1 + (1.0 - m) * s2
    a_b = 2.0 * np.pi / lam * (torch.exp(-lam) - torch.exp(-2.0 * lam))
    a_u = 2.0 * np.pi / lam * (1.0 - torch.exp(-lam))
    return a_b * (1.0 - s) + a_u * s


def lambda_trick(lobe1, lam1, mu1, lobe2, lam2, mu2):
    ratio = lam1 / lam2
    lobe1 = norm_axis(lobe1)
    lobe2 = norm_axis(lobe2)
    dot = torch.sum(lobe1 * lobe2, dim=-1, keepdim=True)
    tmp = torch.sqrt(ratio * ratio + 1.0 + 2.0 * ratio * dot)
    tmp = torch.min(tmp, ratio + 1.0)
    lam3 = lam2 * tmp
    lobes = (ratio / tmp) * lobe1 + (1.0 / tmp) * lobe2
    mus = mu1 * mu2 * torch.exp(lam2 * (tmp - ratio - 1.0))
    return lobes, lam3, mus


# ----------------------------------------------------------------------------------------------------------------------
# visibility sampling (model/sg_render.py:111-301)
# ----------------------------------------------------------------------------------------------------------------------
def _z_axis_like(x):
    z = torch.zeros_like(x)
    z[..., 2] = 1
    return z


def diffuse_sample_dirs(lobes, lambdas, u_theta, u_phi, thr=1.0):
    """sg_render.py:117-146.  lobes [M,3], lambdas [M,1], u_* uniform [M,S].  Returns light_dirs [M,1,3],
    sample_dir [M,S,3]."""
    light_dirs = norm_axis(lobes.unsqueeze(-2))
    lam = lambdas.unsqueeze(-2)
    U = norm_axis(torch.cross(_z_axis_like(light_dirs), light_dirs, dim=-1))
    V = norm_axis(torch.cross(light_dirs, U, dim=-1))
    sharp = torch.clamp(lam[:, :, 0], min=1e-4)
    sg_range = torch.zeros_like(sharp)
    sg_range[:, :] = torch.clamp(sharp.min(), max=thr)
    phi_range = torch.arccos((-0.95 * sg_range) / sharp + 1)
    r_theta = (u_theta * 2 * np.pi).unsqueeze(-1)
    r_phi = (u_phi * phi_range).unsqueeze(-1)
    sample_dir = U * torch.cos(r_theta) * torch.sin(r_phi) + V * torch.sin(r_theta) * torch.sin(r_phi) \
        + light_dirs * torch.cos(r_phi)
    return light_dirs, sample_dir


def get_diffuse_visibility(points, normals, vis_fn, lobes, lambdas, u_theta, u_phi, testing=False,
                           return_aux=False):
    """sg_render.py:111-195 with nsamp = u_theta.shape[1].  vis_fn(points[k,3], dirs[k,3]) -> logits [k,2].
    Returns vis [M, n]."""
    M, S = u_theta.shape
    n = points.shape[0]
    light_dirs, sample_dir = diffuse_sample_dirs(lobes, lambdas, u_theta, u_phi)
    flat = sample_dir.reshape(-1, 3)
    in_dir = flat.unsqueeze(0).expand(n, -1, 3)
    in_p = points.unsqueeze(1).expand(-1, M * S, 3)
    nrm = normals.unsqueeze(1).expand(-1, M * S, 3)
    cos_mask = torch.sum(nrm * in_dir, dim=-1) > TINY
    if testing:
        with torch.no_grad():
            logits = vis_fn(in_p[cos_mask], in_dir[cos_mask])
    else:
        logits = vis_fn(in_p[cos_mask], in_dir[cos_mask])
    pv = torch.softmax(logits, dim=-1)[..., 1]
    vis = torch.zeros(n, M * S)
    vis[cos_mask] = pv
    vis = vis.reshape(n, M, S).permute(1, 2, 0)  # [M,S,n]
    w = torch.exp(lambdas.unsqueeze(-2) * (torch.sum(sample_dir * light_dirs, dim=-1, keepdim=True) - 1.0))
    out = torch.sum(vis * w, dim=1) / (torch.sum(w, dim=1) + TINY)
    if return_aux:
        return out, dict(sample_dir=sample_dir, weight=w, mask=cos_mask, n_query=int(cos_mask.sum()))
    return out


def get_specular_visibility(points, normals, viewdirs, vis_fn, lobes, lambdas, u_theta, u_phi, testing=False,
                            inv=False, return_aux=False):
    """sg_render.py:198-301, single-view branch.  lobes [n,3], lambdas [n,1], u_* uniform [n,S].  Returns vis [n]."""
    S = u_theta.shape[1]
    light_dirs = lobes.unsqueeze(-2)
    lam = lambdas.unsqueeze(-2)
    ndv = torch.clamp(torch.sum(normals * viewdirs, dim=-1, keepdim=True), min=0.0)
    ref = (-viewdirs + 2 * ndv * normals).unsqueeze(-2)
    U = norm_axis(torch.cross(_z_axis_like(ref), ref, dim=-1))
    V = norm_axis(torch.cross(ref, U, dim=-1))
    sharp = torch.clip(lam[..., 0], min=0.1, max=50)
    sg_range = torch.zeros_like(sharp)
    sg_range[:, :] = torch.clamp(sharp.min(), max=1)
    phi_range = torch.arccos((-0.95 * sg_range) / sharp + 1)
    r_theta = (u_theta * 2 * np.pi).unsqueeze(-1)
    r_phi = (u_phi * phi_range).unsqueeze(-1)
    sample_dir = U * torch.cos(r_theta) * torch.sin(r_phi) + V * torch.sin(r_theta) * torch.sin(r_phi) \
        + ref * torch.cos(r_phi)
    in_p = points.unsqueeze(1).expand(-1, S, 3)
    nrm = normals.unsqueeze(1).expand(-1, S, 3)
    cos_mask = torch.sum(nrm * sample_dir, dim=-1) > TINY
    if testing:
        with torch.no_grad():
            logits = vis_fn(in_p[cos_mask], sample_dir[cos_mask])
    else:
        logits = vis_fn(in_p[cos_mask], sample_dir[cos_mask])
    pv = torch.softmax(logits, dim=-1)[..., 0 if inv else 1]
    vis = torch.zeros(points.shape[0], S)
    vis[cos_mask] = pv
    w = torch.exp(sharp * (torch.sum(sample_dir * light_dirs, dim=-1) - 1.0))
    if testing:  # sg_render.py:285-292: rows whose weight sum is inf keep only the inf entries
        bad = torch.isinf(torch.sum(w, dim=-1))
        if bad.any():
            sub = w[bad]
            w = w.clone()
            w[bad] = torch.where(torch.isinf(sub), torch.ones_like(sub), torch.zeros_like(sub))
    out = torch.sum(vis * w, dim=-1) / (torch.sum(w, dim=-1) + TINY)
    if return_aux:
        return out, dict(sample_dir=sample_dir, weight=w, mask=cos_mask, n_query=int(cos_mask.sum()))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# SG renderer (model/sg_render.py:304-565), single-view, metallic=None, fun_spec=False, argmax_vis=False
# ----------------------------------------------------------------------------------------------------------------------
def kl_divergence(x, mu=0.05):
    """utils/utils.py:14-17: KL(mu || mean_0(x)) averaged over the columns (the CESR supervise term)."""
    rho_hat = torch.mean(x, 0)
    rho = torch.full_like(rho_hat, mu)
    return torch.mean(rho * torch.log(rho / (rho_hat + 1e-4)) + (1 - rho) * torch.log((1 - rho) / (1 - rho_hat + 1e-4)))


PREFIT_WEIGHT = {"warmup": 0.1, "project": 0.2}   # sg_render.py:397-403; anything else ("explore", False): 1.0


MU_COS, LAMBDA_COS, ALPHA_COS = 32.7080, 0.0315, 31.7003


def render_with_sg(points, normal, viewdirs, lgtSGs, specular_reflectance, roughness, diffuse_albedo, vis_fn, rnd,
                   comp_vis=True, lin_diff=False, testing=False, indir_integral=None, stats=None, diffuse_vis=None,
                   prefit=False):
    """rnd: dict with 'diff_theta','diff_phi' [M,S] (only comp_vis; S = 32, or 8 when diffuse_vis is given,
    sg_render.py:389) and 'spec_theta','spec_phi' [n,8].  diffuse_vis [n*M] / prefit: the CESR branch
    (sg_render.py:393-407)."""
    M = lgtSGs.shape[1]
    n = normal.shape[0]
    lobes = lgtSGs[..., :3] / (torch.norm(lgtSGs[..., :3], dim=-1, keepdim=True) + TINY)
    lambdas = torch.abs(lgtSGs[..., 3:4])
    mus0 = torch.abs(lgtSGs[..., -3:])
    nrm = normal.unsqueeze(-2).expand(n, M, 3)
    view = viewdirs.unsqueeze(-2).expand(n, M, 3).detach()
    spec_refl = specular_reflectance.unsqueeze(1).expand(n, M, 3)

    vis_shadow = torch.zeros(n, 3)
    supervise = torch.tensor(0.0)
    if comp_vis:
        lv, aux = get_diffuse_visibility(points, nrm[:, 0, :].detach(), vis_fn, lobes[0], lambdas[0],
                                         rnd["diff_theta"], rnd["diff_phi"], testing=testing, return_aux=True)
        if stats is not None:
            stats["n_query"] = stats.get("n_query", 0) + aux["n_query"]
        light_vis_gt = lv.permute(1, 0).unsqueeze(-1).expand(n, M, 3)
        if diffuse_vis is not None:
            assert rnd["diff_theta"].shape[1] == 8
            light_vis = diffuse_vis.reshape(-1, M, 1).expand(n, M, 3)
            if prefit == "warmup":
                supervise = kl_divergence((light_vis_gt.detach() - light_vis).abs()[..., 0], 0.01) * 0.1
                light_vis = light_vis_gt
            else:
                supervise = kl_divergence((light_vis_gt - light_vis).abs()[..., 0], 0.01) * PREFIT_WEIGHT.get(prefit, 1.0)
        else:
            light_vis = light_vis_gt
        vis_shadow = ((light_vis * mus0).sum(1) / torch.clamp(mus0.sum(1), 1e-4)).detach()

    # ---- specular (sg_render.py:414-500)
    inv_r4 = 2.0 / (roughness * roughness * roughness * roughness)
    brdf_lam = inv_r4.unsqueeze(1).expand(n, M, 1)
    brdf_mu = (inv_r4 / np.pi).expand(n, 3).unsqueeze(1).expand(n, M, 3)
    vdl = torch.clamp(torch.sum(nrm * view, dim=-1, keepdim=True), min=0.0)
    wl = 2 * vdl * nrm - view
    wl = wl / (torch.norm(wl, dim=-1, keepdim=True) + TINY)
    wlam = brdf_lam / (4 * vdl + TINY)
    half = wl + view
    half = half / (torch.norm(half, dim=-1, keepdim=True) + TINY)
    vdh = torch.clamp(torch.sum(view * half, dim=-1, keepdim=True), min=0.0)
    fres = spec_refl + (1.0 - spec_refl) * torch.pow(2.0, -(5.55473 * vdh + 6.8316) * vdh)
    d1 = torch.clamp(torch.sum(wl * nrm, dim=-1, keepdim=True), min=0.0)
    d2 = torch.clamp(torch.sum(view * nrm, dim=-1, keepdim=True), min=0.0)
    k = ((roughness + 1.0) * (roughness + 1.0) / 8.0).unsqueeze(1).expand(n, M, 1)
    g1 = d1 / (d1 * (1 - k) + k + TINY)
    g2 = d2 / (d2 * (1 - k) + k + TINY)
    moi = fres * (g1 * g2) / (4 * d1 * d2 + TINY)
    wmu = brdf_mu * moi
    bv, aux = get_specular_visibility(points, nrm[:, 0, :], view[:, 0, :], vis_fn, wl[:, 0], wlam[:, 0],
                                      rnd["spec_theta"], rnd["spec_phi"], testing=testing, inv=not comp_vis,
                                      return_aux=True)
    if stats is not None:
        stats["n_query"] = stats.get("n_query", 0) + aux["n_query"]
    brdf_vis = bv.unsqueeze(-1).unsqueeze(-1).expand(n, M, 3)
    fl, flam, fmu = lambda_trick(lobes, lambdas, mus0 * brdf_vis, wl, wlam, wmu)
    lp, lamp, mup = lambda_trick(nrm, LAMBDA_COS, MU_COS, fl, flam, fmu)
    da = torch.sum(lp * nrm, dim=-1, keepdim=True)
    db = torch.sum(fl * nrm, dim=-1, keepdim=True)
    spec = mup * hemisphere_int(lamp, da) - fmu * ALPHA_COS * hemisphere_int(flam, db)
    spec = torch.clamp(spec.sum(dim=-2), min=0.0)

    # ---- diffuse (sg_render.py:506-536)
    mus = mus0 * light_vis if comp_vis else mus0
    diffuse = (diffuse_albedo / np.pi).unsqueeze(-2).expand(n, M, 3)
    fmu_d = mus if lin_diff else mus * diffuse
    lp, lamp, mup = lambda_trick(nrm, LAMBDA_COS, MU_COS, lobes, lambdas, fmu_d)
    da = torch.sum(lp * nrm, dim=-1, keepdim=True)
    db = torch.sum(lobes * nrm, dim=-1, keepdim=True)
    diff = mup * hemisphere_int(lamp, da) - fmu_d * ALPHA_COS * hemisphere_int(lambdas, db)
    diff = torch.clamp(diff.sum(dim=-2), min=0.0)
    if indir_integral is not None:
        diff = indir_integral if lin_diff else indir_integral * (diffuse_albedo / np.pi)
    return dict(sg_rgb=spec + diff, sg_specular_rgb=spec, sg_diffuse_rgb=diff, vis_shadow=vis_shadow,
                supervise=supervise)


def render_with_all_sg(points, normal, viewdirs, lgtSGs, specular_reflectance, roughness, diffuse_albedo, vis_fn,
                       rnd, indir_integral=None, indir_lgtSGs=None, lin_diff=False, testing=False, stats=None,
                       diffuse_vis=None, prefit=False):
    """sg_render.py:304-337.  rnd keys: diff_theta, diff_phi [M,32]; spec_theta, spec_phi [n,8] (direct pass);
    ind_theta, ind_phi [n,8] (indirect pass) -- the draw order of SURVEY.md A.4."""
    n = normal.shape[0]
    M = lgtSGs.shape[0]
    if lgtSGs.dim() == 2:
        lgtSGs = lgtSGs.unsqueeze(0).expand(n, M, 7)
    ret = render_with_sg(points, normal, viewdirs, lgtSGs, specular_reflectance, roughness, diffuse_albedo, vis_fn,
                         rnd, comp_vis=True, lin_diff=lin_diff, testing=testing, stats=stats, diffuse_vis=diffuse_vis,
                         prefit=prefit)
    ind = dict(indir_rgb=torch.zeros_like(points), indir_diffuse_rgb=torch.zeros_like(points),
               indir_specular_rgb=torch.zeros_like(points))
    if indir_lgtSGs is not None:
        r2 = render_with_sg(points, normal, viewdirs, indir_lgtSGs, specular_reflectance, roughness, diffuse_albedo,
                            vis_fn, dict(spec_theta=rnd["ind_theta"], spec_phi=rnd["ind_phi"]), comp_vis=False,
                            lin_diff=lin_diff, testing=testing, indir_integral=indir_integral, stats=stats)
        ind = dict(indir_rgb=r2["sg_rgb"], indir_diffuse_rgb=r2["sg_diffuse_rgb"],
                   indir_specular_rgb=r2["sg_specular_rgb"])
    ret.update(ind)
    return ret


# ----------------------------------------------------------------------------------------------------------------------
# camera (utils/rend_util.py:51-97,141-163)
# ----------------------------------------------------------------------------------------------------------------------
def quat_to_rot(q):
    """utils/rend_util.py:100-117: rotation matrix of the normalised quaternion (r, i, j, k)."""
    q = F.normalize(q, dim=1)
    qr, qi, qj, qk = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.ones(q.shape[0], 3, 3)
    R[:, 0, 0] = 1 - 2 * (qj ** 2 + qk ** 2)
    R[:, 0, 1] = 2 * (qj * qi - qk * qr)
    R[:, 0, 2] = 2 * (qi * qk + qr * qj)
    R[:, 1, 0] = 2 * (qj * qi + qk * qr)
    R[:, 1, 1] = 1 - 2 * (qi ** 2 + qk ** 2)
    R[:, 1, 2] = 2 * (qj * qk - qi * qr)
    R[:, 2, 0] = 2 * (qk * qi - qj * qr)
    R[:, 2, 1] = 2 * (qj * qk + qi * qr)
    R[:, 2, 2] = 1 - 2 * (qi ** 2 + qj ** 2)
    return R


def camera_rays(uv, pose, intrinsics):
    """get_camera_params + lift (utils/rend_util.py:51-97). uv [B,N,2], pose [B,4,4] or [B,7] (quaternion + position),
    K [B,3,3] -> dirs [B,N,3], cam [B,3]."""
    B, N, _ = uv.shape
    p = torch.eye(4).repeat(B, 1, 1)
    if pose.shape[1] == 7:
        cam_loc = pose[:, 4:]
        p[:, :3, :3] = quat_to_rot(pose[:, :4])
        p[:, :3, 3] = cam_loc
    else:
        cam_loc = pose[:, :3, 3]
        p[:, :3, :4] = pose[:, :3, :4]
    fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
    cx, cy, sk = intrinsics[:, 0, 2], intrinsics[:, 1, 2], intrinsics[:, 0, 1]
    x, y = uv[:, :, 0], uv[:, :, 1]
    z = torch.ones(B, N)
    xl = (x - cx.unsqueeze(-1) + cy.unsqueeze(-1) * sk.unsqueeze(-1) / fy.unsqueeze(-1)
          - sk.unsqueeze(-1) * y / fy.unsqueeze(-1)) / fx.unsqueeze(-1) * z
    yl = (y - cy.unsqueeze(-1)) / fy.unsqueeze(-1) * z
    pc = torch.stack((xl, -yl, -z, torch.ones_like(z)), dim=-1).permute(0, 2, 1)
    world = torch.bmm(p, pc).permute(0, 2, 1)[:, :, :3]
    return F.normalize(world - cam_loc[:, None, :], dim=2), cam_loc


def sphere_intersection(cam_loc, ray_dirs, r=1.0):
    """rend_util.get_sphere_intersection (:141-163) -> [B,N,2] clamped at 0.01, mask [B,N]."""
    B, N, _ = ray_dirs.shape
    b = torch.bmm(ray_dirs, cam_loc.unsqueeze(-1))[..., 0]
    under = (b ** 2 - (cam_loc.norm(2, 1, keepdim=True) ** 2 - r ** 2)).reshape(-1)
    mask = under > 0
    out = torch.zeros(B * N, 2)
    out[mask] = torch.sqrt(under[mask]).unsqueeze(-1) * torch.tensor([-1.0, 1.0])
    out[mask] -= b.reshape(-1)[mask].unsqueeze(-1)
    return out.reshape(B, N, 2).clamp_min(0.01), mask.reshape(B, N)


# ----------------------------------------------------------------------------------------------------------------------
# tone mapping / loss (model/color_correction.py:31-59,112-137; model/loss.py:31-125; train_pbr.py:313-346)
# ----------------------------------------------------------------------------------------------------------------------
def aces(x):
    return x * (2.51 * x + 0.03) / (x * (2.43 * x + 0.59) + 0.14)


def aces_inverse(x):
    return ((0.59 * x - 0.03) + torch.sqrt((0.59 * x - 0.03) ** 2 + 4 * (2.51 - 2.43 * x) * 0.14 * x)) / (
        2 * (2.51 - 2.43 * x))


def hdr_shift_as_input(sd):
    return torch.clamp(sd["gamma.hdr_shift.adapt_illum"] * 10 + 0.5, 0, 1).view(1, 1)


def hdr2ldr(x, shift):
    """hdr_mode=0: aces(x) / clamp(shift,1e-4,1)^0.2."""
    return aces(x) / torch.clamp(shift, 1e-4, 1) ** 0.2


def ldr2hdr(x, shift):
    return aces_inverse(x * torch.clamp(shift, 1e-4, 1) ** 0.2)


def _kl(rho, latent):
    rho_hat = torch.mean(torch.sigmoid(latent), 0)
    rho = torch.full_like(rho_hat, rho)
    return torch.mean(rho * torch.log(rho / (rho_hat + 1e-4)) + (1 - rho) * torch.log((1 - rho) / (1 - rho_hat + 1e-4)))


def pbr_loss(sd, out, rgb_gt, sg_rgb_weight=1.0, kl_weight=1.0, latent_smooth_weight=1.0):
    """InvLoss.forward (loss.py:97-125, L1 sum / N) + PBRTrainRunner.pbr_step/white_loss (train_pbr.py:313-346).
    Weights default to the confs_sg/hotdog.conf:46-58 loss{} values."""
    nm = out["network_object_mask"] & out["object_mask"]
    pred = hdr2ldr(out["sg_rgb"] + out["indir_rgb"], hdr_shift_as_input(sd))
    if nm.sum() == 0:
        rgb_loss = torch.tensor(0.0)
    else:
        rgb_loss = torch.abs(pred[nm] - rgb_gt.reshape(-1, 3)[nm]).sum() / float(out["object_mask"].shape[0])
    smooth = torch.abs(out["diffuse_albedo"] - out["random_xi_diffuse_albedo"]).mean() + \
        torch.abs(out["roughness"][..., 0] - out["random_xi_roughness"][..., 0]).mean() * 0.2
    lat = sparse_ae_encode(sd, MAT + ".spec_brdf_encoder_layer", pe(out["points"][out["network_object_mask"]], 10))
    kl = _kl(0.05, lat)
    lgt = torch.abs(sd[MAT + ".lgtSGs"][..., -3:])
    white = (lgt / (lgt.norm(dim=-1, keepdim=True) + 1e-4)).var(-1).mean() * 0.01
    loss = sg_rgb_weight * rgb_loss + kl_weight * kl * 1.0 + latent_smooth_weight * smooth * 0.1 + white
    return loss, dict(rgb_loss=rgb_loss, kl=kl, smooth=smooth, white=white)


def cesr_loss(sd, out, rgb_gt, cur_iter, smooth_w, kl_w, sg_rgb_weight=1.0, kl_weight=1.0, latent_smooth_weight=1.0):
    """ClusteredAlbedoTrainRunner.pbr_step (training/train_cesr.py:387-430): InvLoss terms only after iteration 500,
    weighted by the explore / project (smooth_w, kl_w) of the conf, plus the hook's supervise term (gradient_error)."""
    loss = torch.tensor(0.0)
    parts = {}
    if cur_iter > 500:
        _, parts = pbr_loss(sd, out, rgb_gt, sg_rgb_weight, kl_weight, latent_smooth_weight)
        loss = sg_rgb_weight * parts["rgb_loss"] + kl_weight * parts["kl"] * kl_w \
            + latent_smooth_weight * parts["smooth"] * smooth_w
    return loss + out["gradient_error"], parts


def illum_loss(out, tr, anneal_t=0.0):
    """IllumLoss.forward + query_indir_illum (model/loss.py:128-179, L1): -> (radiance_loss, visibility_loss)."""
    indir_mask, pm = tr["indir_mask"], out["network_object_mask"]
    sgs = out["indirect_sgs"][pm]
    n, S = tr["sample_dirs"].shape[:2]
    M = sgs.shape[1]
    sg = sgs.unsqueeze(-3).expand(-1, S, -1, -1)
    sdirs = tr["sample_dirs"].unsqueeze(-2).expand(-1, -1, M, -1)
    lobes = sg[..., :3] / torch.norm(sg[..., :3], dim=-1, keepdim=True)
    pred = (sg[..., -3:] * torch.exp(sg[..., 3:4] * (torch.sum(sdirs * lobes, dim=-1, keepdim=True) - 1.))).sum(2)
    gt = tr["trace_radiance"][indir_mask] + anneal_t
    rad = F.l1_loss(gt, pred[indir_mask[pm]]) + F.l1_loss(tr["gt_integral"][pm], out["indir_integral"][pm])
    gt_vis = (~tr["gt_vis"][pm]).long().reshape(-1)
    vis = F.cross_entropy(tr["pred_vis"][pm].reshape(-1, 2), gt_vis)
    return rad, vis
