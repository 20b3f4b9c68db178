"""Full-image rendering in evaluation mode: the compute part of ``PBRTrainRunner.plot_to_disk``
(training/train_pbr.py:235-311) -- ``model.eval()``, ``is_training = False`` (visibility in testing mode), the pixel grid
split into chunks (``utils/general.py:27-38`` uses 1024 pixels; a B200 takes 16 384 and more per call), ``hdr2ldr`` of
the predictions, the per-chunk results merged back into images.  Forward only, nothing is kept for a backward."""
import torch

from ._lib import RobirError


@torch.no_grad()
def render_image(model, pose, intrinsics, H, W, chunk=16384, train_spec=True, out=None):
    """pose [1,4,4] (or [1,7]), intrinsics [1,3,3] on the model's device -> dict of [H*W, C] device tensors with the keys
    ``plot_to_disk`` collects (:271-282): pred_rgb / sg_rgb / indir_rgb (tone-mapped), diffuse_albedo, roughness (expanded to
    3 channels), vis_shadow, plus network_object_mask [H*W].  ``out``: optional dict of preallocated tensors to fill."""
    if chunk < 1:
        raise RobirError("render_image: chunk must be positive")
    dev = pose.device
    was_training, was_is_training = model.training, getattr(model, "is_training", True)
    static = getattr(model, "static_shapes", False)
    model.eval()
    model.is_training = False
    model.static_shapes = False
    try:
        n = H * W
        keys = ("pred_rgb", "sg_rgb", "indir_rgb", "diffuse_albedo", "roughness", "vis_shadow")
        if out is None:
            out = {k: torch.empty(n, 3, device=dev) for k in keys}
            out["network_object_mask"] = torch.empty(n, dtype=torch.bool, device=dev)
        pix = torch.arange(n, device=dev)
        uv = torch.stack([(pix % W).float(), (pix // W).float()], -1)[None]
        ones = torch.ones(1, min(chunk, n), dtype=torch.bool, device=dev)
        tone = model.gamma.hdr_shift
        shift = tone.as_input()
        for lo in range(0, n, chunk):
            hi = min(lo + chunk, n)
            inp = {"uv": uv[:, lo:hi], "object_mask": ones[:, :hi - lo], "pose": pose, "intrinsics": intrinsics,
                   "hdr_shift": shift.expand(hi - lo, 1)}
            o = model(inp, trainstage="Material", lin_diff=False, fun_spec=False, train_spec=train_spec)
            out["sg_rgb"][lo:hi] = tone.hdr2ldr(o["sg_rgb"])
            out["indir_rgb"][lo:hi] = tone.hdr2ldr(o["indir_rgb"])
            out["pred_rgb"][lo:hi] = tone.hdr2ldr(o["sg_rgb"] + o["indir_rgb"])
            out["diffuse_albedo"][lo:hi] = o["diffuse_albedo"]
            out["roughness"][lo:hi] = o["roughness"][..., 0:1].expand(-1, 3)
            out["vis_shadow"][lo:hi] = o["vis_shadow"]
            out["network_object_mask"][lo:hi] = o["network_object_mask"]
        return out
    finally:
        model.train(was_training)
        model.is_training = was_is_training
        model.static_shapes = static
