"""Diagnostic: kernel timeline (start, duration, stream) of one replay of the graphed PBR step, via torch.profiler/CUPTI.
Writes gpurun_out/timeline.csv.  Not a bench: profiler overhead is inside the numbers."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robir_b200  # noqa: E402
from robir_b200 import rng, synthetic  # noqa: E402
from robir_b200.graph import GraphedPBRStep  # noqa: E402
from robir_b200.loss import InvLoss  # noqa: E402


def main():
    N, M = 1024, 128
    dev = torch.device("cuda")
    rng.set_mode("device")
    sd = synthetic.synthetic_state_dict(0, num_lgt_sgs=M, sdf_radius=0.87)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=M)))
    model.load_state_dict(sd, strict=True)
    model.to(dev).train()
    model.generate()
    params = list(model.gamma.parameters()) + list(model.envmap_material_network.parameters())
    opt = torch.optim.Adam(params, lr=5e-4, capturable=True, fused=True)
    pose, K = synthetic.camera_pose().to(dev), synthetic.camera_intrinsics().to(dev)
    pipe = os.environ.get("ROBIR_PIPELINE_TRACE", "1") == "1"       # what bench.py runs by default
    step = GraphedPBRStep(model, InvLoss(), opt, N, pose, K, pipeline_trace=pipe)
    pix = synthetic.training_pixels(0, n=N)
    uv = torch.stack([(pix % 800).float(), (pix // 800).float()], -1)[None].to(dev)
    om = torch.ones(1, N, dtype=torch.bool, device=dev)
    gt = torch.full((1, N, 3), 0.5, device=dev)
    args = (uv, om, gt, uv, om) if pipe else (uv, om, gt)
    for _ in range(3):
        step(*args)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            step(*args)
        torch.cuda.synchronize()
    rows = []
    try:                                     # kineto events carry the stream id (device_resource_id)
        for ev in prof.profiler.kineto_results.events():
            if ev.device_type() == torch.autograd.DeviceType.CUDA:
                rows.append((ev.start_ns() / 1e3, ev.duration_ns() / 1e3, ev.name(), ev.device_resource_id()))
    except Exception:
        rows = []
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, ev.name, -1))
    rows.sort()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "timeline.csv"), "w") as f:
        for s, d, n, st in rows:
            f.write("%.3f,%.3f,%s,%d\n" % (s, d, n.replace(",", ";")[:120], st))
    print("kernels:", len(rows))


if __name__ == "__main__":
    main()
