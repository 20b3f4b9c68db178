"""Seed-reproducible synthetic weights and camera workloads (no dataset / NeuS checkpoint exists offline).

``synthetic_state_dict`` produces the 134 tensors of the reference ``IDRNetwork.state_dict()`` (key names and shapes
verified against the reference, SURVEY.md section 5 "Checkpoint / resume") from a CPU generator, so tests, goldens,
``bench.py`` and the reference all see the same weights without shipping 21 MB files:

  * SDF network: IDR geometric initialisation (sphere of radius ``sdf_radius`` in NeuS coordinates, i.e. half that in
    stage-2 coordinates) following the scheme of model/neus_model.py:358-376, optionally perturbed;
  * colour / material / indirect / visibility networks: uniform(+-1/sqrt(fan_in)) like nn.Linear's default;
    visibility hidden weights can be scaled (``vis_gain``) so the classifier's logits are not all ~0;
  * light SGs: fibonacci lobes with log-uniform sharpness in [1.2, 500] (the range of the fitted envmaps shipped with
    the reference) and energy-normalised amplitudes.

``camera_inputs`` builds the blender-style 800x800 camera of BASELINE.md section 3 (off-axis so no ray hits the octree
NaN edge case).
"""
import math

import torch

SDF = "implicit_network.neus_model.sdf_network"
COL = "implicit_network.neus_model.color_network"


def _float32_default(fn):
    """The weights are defined by the float32 random stream: a caller that switched torch's default dtype (the float64
    oracle pins) must still get the same tensors."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float32)
        try:
            return fn(*a, **k)
        finally:
            torch.set_default_dtype(old)
    return wrapped


def _uniform_linear(gen, out_f, in_f):
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def _put_wn(sd, key, w, b):
    sd[key + ".bias"] = b
    sd[key + ".weight_g"] = w.norm(dim=1, keepdim=True)
    sd[key + ".weight_v"] = w


def _put_seq(sd, prefix, gen, dims):
    for i in range(len(dims) - 1):
        w, b = _uniform_linear(gen, dims[i + 1], dims[i])
        sd["%s.%d.weight" % (prefix, 2 * i)] = w
        sd["%s.%d.bias" % (prefix, 2 * i)] = b


def _put_sparse_ae(sd, prefix, gen, in_dim, out_dim):
    _put_seq(sd, prefix + ".brdf_encoder_layer", gen, [in_dim, 512, 512, 512, 512, 32])
    _put_seq(sd, prefix + ".brdf_decoder_layer", gen, [32, 128, 128, out_dim])


def fibonacci_lobes(k):
    i = torch.arange(k, dtype=torch.float64)
    y = 1 - (i / max(k - 1, 1)) * 2
    r = torch.sqrt(torch.clamp(1 - y * y, min=0))
    th = math.pi * (3.0 - math.sqrt(5.0)) * i
    return torch.stack([torch.cos(th) * r, y, torch.sin(th) * r], -1).float()


def synthetic_light_sgs(gen, M):
    sg = torch.zeros(M, 7)
    half = max(M // 2, 1)
    lob = fibonacci_lobes(half)
    sg[:half, :3] = lob
    sg[half:, :3] = lob[: M - half]
    sg[:, :3] *= 0.3 + 3.0 * torch.rand(M, 1, generator=gen)          # un-normalised lobes, like the fitted envmaps
    sg[:, 3] = torch.exp(math.log(1.2) + torch.rand(M, generator=gen) * (math.log(500.0) - math.log(1.2)))
    mu = torch.rand(M, 1, generator=gen).expand(M, 3) * (0.7 + 0.6 * torch.rand(M, 3, generator=gen))
    lam = sg[:, 3:4]
    energy = (mu * 2.0 * math.pi / lam * (1.0 - torch.exp(-2.0 * lam))).sum(0, keepdim=True)
    sg[:, 4:] = mu / energy * 2.0 * math.pi * 0.8
    return sg


@_float32_default
def synthetic_state_dict(seed=0, num_lgt_sgs=128, sdf_radius=0.5, perturb=0.0, vis_gain=2.0):
    """sdf_radius is the geometric-init bias; because softplus(0) != 0 the zero level set sits at a stage-2 radius of
    about 0.33 for 0.5 (small octree, used by tests/goldens) and about 0.60 for 0.87 (bench workload)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    # ---- SDF network (NeuS coordinates): 63 -> 256 x3 -> 193 (+63 skip) -> 256 x4 -> 257
    dims = [63] + [256] * 8 + [257]
    for l in range(9):
        out_dim = dims[l + 1] - dims[0] if l + 1 == 4 else dims[l + 1]
        in_dim = dims[l]
        if l == 8:
            w = math.sqrt(math.pi) / math.sqrt(in_dim) + 1e-4 * torch.randn(out_dim, in_dim, generator=gen)
            b = torch.full((out_dim,), -float(sdf_radius))
        else:
            w = torch.randn(out_dim, in_dim, generator=gen) * (math.sqrt(2) / math.sqrt(out_dim))
            b = torch.zeros(out_dim)
            if l == 0:
                w[:, 3:] = 0.0
            elif l == 4:
                w[:, -(dims[0] - 3):] = 0.0
        if perturb > 0:
            w = w + perturb * torch.randn(out_dim, in_dim, generator=gen) * w.abs().mean()
        _put_wn(sd, "%s.lin%d" % (SDF, l), w, b)
    # ---- colour network 289 -> 256 x4 -> 3
    cd = [289, 256, 256, 256, 256, 3]
    for l in range(5):
        w, b = _uniform_linear(gen, cd[l + 1], cd[l])
        _put_wn(sd, "%s.lin%d" % (COL, l), w, b)
    sd["implicit_network.neus_model.deviation_network.variance"] = torch.tensor(0.3)
    # ---- indirect illumination
    _put_seq(sd, "indirect_illum_network.lobe_layer", gen, [64, 512, 512, 512, 512, 144])
    _put_sparse_ae(sd, "indirect_illum_network.integral_layer", gen, 64, 3)
    # ---- visibility network 126 -> 256 x4 -> 2
    _put_seq(sd, "visibility_network.vis_layer", gen, [126, 256, 256, 256, 256, 2])
    for i in (2, 4, 6, 8):
        sd["visibility_network.vis_layer.%d.weight" % i] *= vis_gain
    # ---- material network
    spec = torch.zeros(1, 1)
    spec[:] = 0.05
    sd["envmap_material_network.specular_reflectance"] = spec
    sd["envmap_material_network.lgtSGs"] = synthetic_light_sgs(gen, num_lgt_sgs)
    _put_sparse_ae(sd, "envmap_material_network.brdf_encoder_layer", gen, 63, 5)
    _put_sparse_ae(sd, "envmap_material_network.spec_brdf_encoder_layer", gen, 63, 5)
    _put_sparse_ae(sd, "envmap_material_network.normal_decoder_layer", gen, 60, 3)
    # ---- tone mapping scalars (model/color_correction.py:9-19,78)
    sd["gamma.gamma"] = torch.tensor(1.0)
    sd["gamma.indir_coef"] = torch.tensor(1.0)
    sd["gamma.dir_coef"] = torch.tensor(2.0)
    sd["gamma.coef"] = torch.tensor(1.0)
    sd["gamma.hdr_shift.adapt_illum"] = torch.tensor(0.0)
    return sd


@_float32_default
def cesr_state_dicts(seed=0, gain=2.0):
    """Seeded weights of the CESR stage's two extra weight-normed MLPs (training/train_cesr.py:106-110; state-dict keys
    of the reference SDFNetwork: ``lin{l}.weight_g | weight_v | bias``): shadow_net 191 -> 512 x8 -> 2 and normal_net
    63 -> 512 x8 -> 3, skip concat at layer 4.  He-style hidden layers; the output layer is scaled by ``gain`` so the
    shadow classifier is not stuck at 0.5 (the reference's geometric init makes both logits equal).  The columns that
    multiply PE band k are attenuated by 2^-k (the spectral decay of a trained network): without it one ulp of a traced
    hit point moves the outputs by 3e-5 and the parity tests would measure the tracer, not these networks."""
    gen = torch.Generator().manual_seed(seed + 7919)
    band = torch.ones(63)
    for k in range(10):
        band[3 + 6 * k: 9 + 6 * k] = 2.0 ** -k
    out = []
    for d_in, d_out in ((191, 2), (63, 3)):
        dims = [d_in] + [512] * 8 + [d_out]
        sd = {}
        for l in range(9):
            o = dims[l + 1] - dims[0] if l + 1 == 4 else dims[l + 1]
            std = math.sqrt(2) / math.sqrt(o) if l < 8 else gain / math.sqrt(dims[l])
            w = torch.randn(o, dims[l], generator=gen) * std
            if l == 0:
                w[:, :63] *= band
            elif l == 4:
                w[:, dims[l] - d_in: dims[l] - d_in + 63] *= band
            b = (torch.rand(o, generator=gen) * 2 - 1) * 0.05
            if l == 8 and d_out == 3:
                b = b + torch.tensor([0.2, -0.3, 2.0])   # normals that mostly face the +z camera of camera_pose(), so that
                #                                          the specular lobe (and its normal gradient) is exercised
            _put_wn(sd, "lin%d" % l, w, b)
            sd["lin%d.weight_g" % l] = sd["lin%d.weight_g" % l] * (0.9 + 0.2 * torch.rand(o, 1, generator=gen))
        out.append(sd)
    return out[0], out[1]


def neus_checkpoint_from(sd):
    """The stage-1 ``model`` dict (NeuSModel.state_dict() keys) contained in a stage-2 state dict."""
    pre = "implicit_network.neus_model."
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


def camera_pose():
    pose = torch.eye(4)[None].clone()
    pose[0, :3, 3] = torch.tensor([0.013, 0.021, 2.0])
    return pose


def camera_intrinsics(H=800, W=800, focal=1111.1):
    return torch.tensor([[focal, 0.0, W / 2 - 0.37], [0.0, focal, H / 2 - 0.21], [0.0, 0.0, 1.0]])[None]


def camera_inputs(pixel_idx, H=800, W=800, focal=1111.1, hdr_shift=0.5):
    """pixel_idx: LongTensor [N] of raster indices into an HxW image -> the dict IDRNetwork.forward consumes
    (uv [1,N,2] as (x,y) float, pose [1,4,4], intrinsics [1,3,3], object_mask [1,N], hdr_shift [N,1])."""
    n = pixel_idx.shape[0]
    uv = torch.stack([(pixel_idx % W).float(), (pixel_idx // W).float()], -1)[None]
    return dict(uv=uv, pose=camera_pose(), intrinsics=camera_intrinsics(H, W, focal),
                object_mask=torch.ones(1, n, dtype=torch.bool), hdr_shift=torch.full((n, 1), float(hdr_shift)))


def training_pixels(step, n=1024, H=800, W=800, seed=2024, crop=None):
    """The PBR stage samples ``num_pixels`` random pixels of one image per iteration (datasets/syn_dataset.py:167-171,
    hotdog.conf:9); here a seeded permutation per step.  ``crop`` restricts to a centred crop x crop window."""
    gen = torch.Generator().manual_seed(seed + step)
    if crop is None:
        return torch.randperm(H * W, generator=gen)[:n]
    sel = torch.randperm(crop * crop, generator=gen)[:n]
    y = sel // crop + (H - crop) // 2
    x = sel % crop + (W - crop) // 2
    return y * W + x
