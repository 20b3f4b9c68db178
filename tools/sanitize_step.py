"""One small PBR training step (M = 16 light SGs, 400 rays, fixed-capacity forward + loss + backward) for
compute-sanitizer: exercises vis_tc_kernel (forward / backward, several tiles per CTA), tc_layer_kernel, wgrad_kernel,
octree_cast_kernel (cooperative lock-step walk) and the fused loss.

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


def main():
    torch.manual_seed(0)
    M, N = 16, int(os.environ.get("ROBIR_SAN_RAYS", "400"))
    sd = synthetic.synthetic_state_dict(0, num_lgt_sgs=M)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=M)))
    model.load_state_dict(sd, strict=True)
    model.cuda().train()
    model.generate()
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
