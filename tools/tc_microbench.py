"""Diagnostic: times robir_vis_tc_fwd / bwd alone on a synthetic pair list (148 x 40 tiles by default) and prints the
time per tile round and the algorithmic TFLOP/s.  Not a bench (no clocks record, no L2 flush)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robir_b200 import ops, synthetic  # noqa: E402
from robir_b200._lib import check, lib, ptr, sm_count, stream  # noqa: E402
import robir_b200  # noqa: E402


def main():
    tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 40
    dev = torch.device("cuda")
    sd = synthetic.synthetic_state_dict(0, num_lgt_sgs=16)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=16)))
    model.load_state_dict(sd, strict=True)
    model.to(dev)
    from robir_b200 import sg_render
    W = sg_render._weights_of(model.visibility_network).get()
    n_pts, n_dirs = 600, 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    tabA = torch.randn(n_pts, 256, device=dev, generator=g)
    rows = tiles * 128
    rowA = (torch.arange(rows, device=dev) // 2048 % n_pts).int()
    rowB = torch.randint(0, n_dirs, (rows,), device=dev, generator=g).int()
    n_tiles = torch.tensor([tiles], dtype=torch.int32, device=dev)
    vis = torch.empty(rows, device=dev)
    mask = torch.empty(rows, 4, 8, dtype=torch.int32, device=dev)
    g_vis = torch.randn(rows, device=dev, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(n_dirs, 3, device=dev, generator=g), dim=-1)
    g_dirs = torch.zeros(n_dirs, 3, device=dev)
    tabB = torch.randn(n_dirs, 256, device=dev, generator=g)

    def fwd():
        check(lib().robir_vis_tc_fwd(ptr(tabA), ptr(tabB), ptr(rowA), ptr(rowB), ptr(n_tiles), tiles, ptr(W["tc_fwd"]),
                                     ptr(W["bias3"]), ptr(W["wd"]), ptr(W["bd"]), ptr(vis), ptr(mask), sm_count(), stream()))

    def bwd():
        check(lib().robir_vis_tc_bwd(ptr(rowB), ptr(n_tiles), tiles, ptr(W["tc_bwd"]), ptr(W["wd"]), ptr(vis), ptr(g_vis),
                                     ptr(mask), ptr(dirs), ptr(g_dirs), sm_count(), stream()))

    flop = 458752.0 * rows
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("%s  %.3f ms  %.2f us/tile-round  %.1f TFLOP/s(alg)" % (
            name, ms, 1e3 * ms / (tiles / 148.0), flop / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
