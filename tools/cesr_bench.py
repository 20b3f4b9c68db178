"""CESR-stage step (SURVEY.md section 8c config 4 inputs: the PBR workload of bench.py + shadow_net / normal_net,
explore phase, lin_diff, S = 8) timed on one GPU, eager dynamic-shape forward.  Diagnostic, not the bench contract:

    python tools/cesr_bench.py [--steps 10] [--warmup 3] [--engine tc|torch] [--json out.json]

Prints ms/step (CUDA events), rays/s, and the time of the shadow_net chain alone (forward, forward + backward) on the
step's n_hit * 128 rows, for the selected engine of the weight-normed chains.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import robir_b200  # noqa: E402
from robir_b200 import cesr, ops, rng, synthetic  # noqa: E402
from robir_b200.loss import InvLoss  # noqa: E402

N_RAYS, M_LOBES, SDF_RADIUS, SEED = 1024, 128, 0.87, 0


def timed(fn, reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--engine", default="tc", choices=["tc", "torch"])
    ap.add_argument("--json", default=None)
    ap.add_argument("--profile", default=None, help="write a per-kernel table of 2 steps (torch.profiler) to this file")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ops.ENGINE["wn"] = args.engine
    rng.set_mode("device")
    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=M_LOBES, sdf_radius=SDF_RADIUS)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=M_LOBES)))
    model.load_state_dict(sd, strict=True)
    model.to(dev).train()
    model.generate()
    model.static_shapes = False
    sh, nr = synthetic.cesr_state_dicts(SEED)
    shadow, normal = cesr.WnMLP(191, 2), cesr.WnMLP(63, 3)
    shadow.load_state_dict(sh)
    normal.load_state_dict(nr)
    hook = cesr.ClusteredAlbedoHook(model, shadow.to(dev), normal.to(dev), cur_iter=600)     # explore phase
    model.get_sg_render = hook.get_sg_render
    loss_fn = InvLoss()
    params = list(model.gamma.parameters()) + list(model.envmap_material_network.parameters()) + hook.parameters()
    opt = torch.optim.Adam(params, lr=5e-4)                                                  # train_cesr.py:111-117
    pose, K = synthetic.camera_pose().to(dev), synthetic.camera_intrinsics().to(dev)
    state = {"step": 0, "hits": 0, "loss": 0.0}

    def step():
        pix = synthetic.training_pixels(state["step"], n=N_RAYS).to(dev)
        uv = torch.stack([(pix % 800).float(), (pix // 800).float()], -1)[None]
        inp = {"uv": uv, "object_mask": torch.ones(1, N_RAYS, dtype=torch.bool, device=dev), "pose": pose,
               "intrinsics": K, "hdr_shift": model.gamma.hdr_shift.as_input().expand(N_RAYS, 1)}
        out = model(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss, _ = hook.pbr_step(loss_fn, out, {"rgb": torch.full((1, N_RAYS, 3), 0.5, device=dev)})
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        state["step"] += 1
        state["hits"] = int(out["network_object_mask"].sum())
        state["loss"] = float(loss)

    for _ in range(args.warmup):
        step()
    ms = timed(step, args.steps)

    # the shadow_net chain alone on the same row count
    rows = max(state["hits"], 1) * M_LOBES
    emb = torch.randn(rows // M_LOBES, 63, device=dev) * 0.3

    def chain_fwd():
        with torch.no_grad():
            cesr.shadow_logits(shadow, emb, M_LOBES)

    def chain_fwd_bwd():
        shadow.zero_grad(set_to_none=True)
        cesr.shadow_logits(shadow, emb, M_LOBES).sum().backward()

    chain_fwd_bwd()
    res = {"engine": args.engine, "ms_per_step": ms, "rays_per_s": N_RAYS / ms * 1e3, "hits": state["hits"],
           "rows": rows, "loss": state["loss"], "shadow_fwd_ms": timed(chain_fwd, 5),
           "shadow_fwd_bwd_ms": timed(chain_fwd_bwd, 5),
           "shadow_fwd_gflop": 2.0 * rows * (191 * 512 + 2 * 512 * 512 + 512 * 321 + 4 * 512 * 512 + 512 * 2) / 1e9}
    res["shadow_fwd_tflops"] = res["shadow_fwd_gflop"] / res["shadow_fwd_ms"]
    print(json.dumps(res))
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            step()
            torch.cuda.synchronize()
        with open(args.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
