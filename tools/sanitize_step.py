"""compute-sanitizer targets (ROBIR_SAN_PARTS = comma list, default all):

  pbr    one small PBR training step (M = 16 light SGs, 400 rays; dynamic and fixed-capacity forward + loss + backward):
         vis_tc_kernel (forward / backward, several tiles per CTA), tc_layer_kernel, wgrad_kernel, octree_cast_kernel
         (cooperative lock-step walk), the fused loss
  sdf    sdf_tc_kernel (tensor-core SDF network): value only, value + normal jets, + features, on 2 500 points
  cesr   the weight-normed chain on 4 500 rows, forward + backward: tc_layer_big_kernel (persistent 128x256 tiles),
         tl_pack_rows_t_kernel / tl_wgrad_kernel / tl_wgrad_reduce_kernel, with and without a device-side active-row count
  neus   the stage-1 NeuS renderer kernels (upsample / merge / midpoints / composite) on 48 rays

    compute-sanitizer --tool memcheck  python tools/sanitize_step.py
    compute-sanitizer --tool racecheck python tools/sanitize_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import robir_b200  # noqa: E402
from robir_b200 import ops, rng, synthetic  # noqa: E402
from robir_b200.loss import InvLoss, pbr_step_loss  # noqa: E402


def part_sdf(model):
    gen = torch.Generator().manual_seed(1)
    pts = (torch.randn(2500, 3, generator=gen) * 0.4).cuda()
    net = model.implicit_network
    old = ops.ENGINE["sdf"]
    for eng in ("tc", "ffma"):
        ops.ENGINE["sdf"] = eng
        with torch.no_grad():
            sdf = net.sdf(pts)
            grad = net.gradient(pts)[:, 0, :]
            out = net(pts)
        torch.cuda.synchronize()
        print("sdf engine %s: |sdf| %.4f |grad| %.4f out %s" % (eng, float(sdf.abs().mean()), float(grad.norm(dim=-1).mean()),
                                                               tuple(out.shape)), flush=True)
    ops.ENGINE["sdf"] = old


def part_cesr():
    from robir_b200 import cesr
    sh, _ = synthetic.cesr_state_dicts(0)
    net = cesr.WnMLP(191, 2)
    net.load_state_dict(sh)
    net.cuda()
    gen = torch.Generator().manual_seed(2)
    rows = 4500
    x = (torch.randn(rows, 191, generator=gen) * 0.4).cuda()
    gup = torch.randn(rows, 2, generator=gen).cuda()
    for n_active in (None, 1536):
        na = None if n_active is None else torch.tensor([n_active], dtype=torch.int32, device="cuda")
        g = gup.clone()
        if n_active is not None:
            g[n_active:] = 0
        net.zero_grad()
        out = cesr.wn_mlp(net, x, n_active=na)
        (out * g).sum().backward()
        torch.cuda.synchronize()
        print("wn chain rows %d n_active %s: |out| %.4f |d lin0.weight_v| %.3e" % (
            rows, n_active, float(out.abs().mean()), float(net.lin0.weight_v.grad.abs().max())), flush=True)


def part_neus(model):
    from robir_b200 import neus_stage1
    gen = torch.Generator().manual_seed(3)
    B = 48
    o = torch.nn.functional.normalize(torch.randn(B, 3, generator=gen), dim=-1) * 3.0
    d = torch.nn.functional.normalize(-o + 0.4 * torch.randn(B, 3, generator=gen), dim=-1)
    near, far = torch.full((B, 1), 1.0), torch.full((B, 1), 5.0)
    with torch.no_grad():
        r = neus_stage1.render_neus(model.implicit_network, o.cuda(), d.cuda(), near.cuda(), far.cuda(),
                                    torch.rand(B, 1, generator=gen).cuda(), 1.0)
    torch.cuda.synchronize()
    print("neus stage 1: acc %.4f" % float(r["acc"].mean()), flush=True)


def main():
    torch.manual_seed(0)
    parts = os.environ.get("ROBIR_SAN_PARTS", "pbr,sdf,cesr,neus").split(",")
    M, N = 16, int(os.environ.get("ROBIR_SAN_RAYS", "400"))
    sd = synthetic.synthetic_state_dict(0, num_lgt_sgs=M)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=M)))
    model.load_state_dict(sd, strict=True)
    model.cuda().train()
    model.generate()
    if "sdf" in parts:
        part_sdf(model)
    if "cesr" in parts:
        part_cesr()
    if "neus" in parts:
        part_neus(model)
    if "pbr" not in parts:
        return
    rng.set_mode("device")
    loss_fn = InvLoss()
    for static in (False, True):
        model.static_shapes = loss_fn.static_shapes = static
        inp = {k: v.cuda() for k, v in synthetic.camera_inputs(synthetic.training_pixels(1, n=N, crop=300)).items()}
        inp["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(N, 1)
        out = model(inp, trainstage="Material", train_spec=True)
        loss, _ = pbr_step_loss(model, loss_fn, out, {"rgb": torch.full((1, N, 3), 0.5).cuda()})
        model.zero_grad()
        loss.backward()
        torch.cuda.synchronize()
        print("static=%s hits %d/%d loss %.6f pairs %d |d lgtSGs| %.3e" % (
            static, int(out["network_object_mask"].sum()), N, float(loss), int(ops.Stats.total()),
            float(model.envmap_material_network.lgtSGs.grad.abs().max())), flush=True)


if __name__ == "__main__":
    main()
