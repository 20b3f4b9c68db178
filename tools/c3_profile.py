"""Diagnostic: per-kernel table (torch.profiler) of two Vis-stage iterations (bench.py --config c3 workload), and the
SDF tensor-core kernel alone on the borrow_color batch size (value + normal + features)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from robir_b200 import ops, synthetic  # noqa: E402
import bench_configs as BC  # noqa: E402
from robir_b200.loss import IllumLoss  # noqa: E402


def main():
    N, NS, M = 256, 512, 128
    rank, world, local, dev, sd, model = BC._setup(M)
    model.indirect_illum_network.train_weights = True
    illum_loss = IllumLoss()
    illum_opt = torch.optim.Adam(model.indirect_illum_network.parameters(), lr=5e-4)
    vis_opt = torch.optim.Adam(model.visibility_network.parameters(), lr=5e-4)

    def step(s):
        inp = {k: v.to(dev) for k, v in synthetic.camera_inputs(synthetic.training_pixels(s, n=N)).items()}
        inp["hdr_shift"] = torch.rand(N, 1, device=dev)
        out = model(inp, trainstage="Illum")
        tr = model.trace_radiance(out, nsamp=NS)
        rad, vis = illum_loss(out, tr, 0.0)
        vis_opt.zero_grad()
        vis.backward(retain_graph=True)
        vis_opt.step()
        illum_opt.zero_grad()
        rad.backward()
        illum_opt.step()
    for s in range(3):
        step(s)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(3)
        step(4)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
    # ---- the SDF kernel alone
    w = model.implicit_network._w
    ops.SDF_TC_MIN_ROWS = 1
    for n, jet, feat in ((740000, True, True), (740000, False, False), (550, True, False)):
        pts = (torch.rand(n, 3, device=dev) * 2 - 1) * 0.7
        for eng in ("tc", "ffma"):
            ops.ENGINE["sdf"] = eng
            for _ in range(2):
                ops.sdf_eval(w, pts, want_grad=jet, want_feat=feat)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.sdf_eval(w, pts, want_grad=jet, want_feat=feat)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            rows = n * (4 if jet else 1)
            flop = 2.0 * rows * (63 * 256 + 6 * 256 * 256 + 256 * 193 + 256 + (256 * 256 if feat else 0))
            print("sdf_eval n=%d jet=%s feat=%s engine=%s: %.3f ms, %.1f TFLOP/s algorithmic (rows x layers)" % (
                n, jet, feat, eng, ms, flop / ms / 1e9), flush=True)
    ops.ENGINE["sdf"] = "tc"


if __name__ == "__main__":
    main()
