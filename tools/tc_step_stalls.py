"""Diagnostic: per-CTA clock spread and stall breakdown of the visibility kernels inside a REAL bench step (eager
fixed-capacity forward + loss + backward at BASELINE config 2), read from the kernel's debug buffer after the forward
(last launch = BRDF-lobe list... the diffuse launch is read through a hook) and after the backward."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robir_b200  # noqa: E402
from robir_b200 import ops, rng, synthetic  # noqa: E402
from robir_b200._lib import check, lib, ptr, sm_count  # noqa: E402
from robir_b200.loss import InvLoss, pbr_step_loss  # noqa: E402


def report(name, dbg):
    d = dbg.double().cpu()
    d = d[d[:, 0] > 0]
    tot = d[:, 0]
    print("%-12s CTAs %3d  issuer clocks min %.0f mean %.0f max %.0f (max/mean %.2f)  waits: weights %.0f%% A %.0f%% "
          "Dfree %.0f%% | epilogue accum %.0f%%" % (name, d.shape[0], tot.min(), tot.mean(), tot.max(),
                                                   tot.max() / tot.mean(), 100 * (d[:, 1] / tot).mean(),
                                                   100 * (d[:, 2] / tot).mean(), 100 * (d[:, 3] / tot).mean(),
                                                   100 * (d[:, 5] / d[:, 4].clamp(min=1)).mean()), flush=True)


def main():
    N, M = 1024, 128
    dev = torch.device("cuda")
    rng.set_mode("device")
    sd = synthetic.synthetic_state_dict(0, num_lgt_sgs=M, sdf_radius=0.87)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=M)))
    model.load_state_dict(sd, strict=True)
    model.to(dev).train()
    model.generate()
    model.static_shapes = True
    loss_fn = InvLoss()
    loss_fn.static_shapes = True
    inp = {k: v.to(dev) for k, v in synthetic.camera_inputs(synthetic.training_pixels(0, n=N)).items()}
    gt = torch.full((1, N, 3), 0.5, device=dev)
    dbg = torch.zeros(sm_count(), 8, dtype=torch.int64, device=dev)
    # hook the two big launches: copy the buffer right after each
    snaps = {}
    orig_fwd, orig_bwd = ops._vis_mlp_fwd, ops._vis_mlp_bwd

    def fwd(W, tabA, tabB, rowA, rowB, n_tiles, max_tiles, need_mask):
        out = orig_fwd(W, tabA, tabB, rowA, rowB, n_tiles, max_tiles, need_mask)
        if max_tiles > 2048:
            snaps["fwd"] = dbg.clone()
        return out

    def bwd(W, rowB, n_tiles, max_tiles, vis, g_vis, mask, dirs, engine):
        out = orig_bwd(W, rowB, n_tiles, max_tiles, vis, g_vis, mask, dirs, engine)
        if max_tiles > 2048:
            snaps["bwd"] = dbg.clone()
        return out
    ops._vis_mlp_fwd, ops._vis_mlp_bwd = fwd, bwd
    for it in range(3):
        if it == 2:
            check(lib().robir_tc_debug_buffer(ptr(dbg)))
        inp["hdr_shift"] = model.gamma.hdr_shift.as_input().expand(N, 1)
        out = model(inp, trainstage="Material", train_spec=True)
        loss, _ = pbr_step_loss(model, loss_fn, out, {"rgb": gt})
        model.zero_grad()
        loss.backward()
        torch.cuda.synchronize()
    check(lib().robir_tc_debug_buffer(None))
    for k in ("fwd", "bwd"):
        report(k, snaps[k])


if __name__ == "__main__":
    main()
