"""Diagnostic: times robir_vis_tc_fwd / bwd alone on a synthetic pair list (148 x 60 tiles by default = the bench step's
diffuse launch) for both engines (terms = 3 parity, 1 fast) and prints the time per tile round, the algorithmic TFLOP/s
and the per-role stall breakdown the kernel records (robir_tc_debug_buffer): which hand-off the MMA issuer, the epilogue
and the weight producer wait on, in % of the kernel's duration (mean over the CTAs).

    python tools/tc_microbench.py [tiles] [--json out.json]

Not a bench (no clocks record, no L2 flush)."""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robir_b200 import ops, synthetic  # noqa: E402
from robir_b200._lib import check, lib, ptr, sm_count, stream  # noqa: E402
import robir_b200  # noqa: E402


def main():
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    args = [a for a in sys.argv[1:] if not a.startswith("--") and a != out_json]
    tiles = int(args[0]) if args else 148 * 60
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
    dbg = torch.zeros(sm_count(), 8, dtype=torch.int64, device=dev)
    results = {}
    for terms in (3, 1):
        def fwd():
            check(lib().robir_vis_tc_fwd(ptr(tabA), ptr(tabB), ptr(rowA), ptr(rowB), ptr(n_tiles), tiles,
                                         ptr(W["tc_fwd%d" % terms]), ptr(W["bias3"]), ptr(W["wd"]), ptr(W["bd"]),
                                         ptr(vis), ptr(mask), terms, sm_count(), stream()))

        def bwd():
            check(lib().robir_vis_tc_bwd(ptr(rowB), ptr(n_tiles), tiles, ptr(W["tc_bwd%d" % terms]), ptr(W["wd"]),
                                         ptr(vis), ptr(g_vis), ptr(mask), ptr(dirs), ptr(g_dirs), terms, sm_count(),
                                         stream()))

        flop = 458752.0 * rows
        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            check(lib().robir_tc_debug_buffer(None))
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
            check(lib().robir_tc_debug_buffer(ptr(dbg)))
            dbg.zero_()
            fn()
            torch.cuda.synchronize()
            check(lib().robir_tc_debug_buffer(None))
            d = dbg.double().cpu()
            tot_i, tot_e = d[:, 0].clamp(min=1), d[:, 4].clamp(min=1)
            st = {"issuer_wait_weights": (d[:, 1] / tot_i).mean().item(), "issuer_wait_A": (d[:, 2] / tot_i).mean().item(),
                  "issuer_wait_D_drained": (d[:, 3] / tot_i).mean().item(),
                  "epilogue_wait_accum": (d[:, 5] / tot_e).mean().item(),
                  "epilogue_wait_X_free": (d[:, 6] / tot_e).mean().item(),
                  "producer_wait_slot": (d[:, 7] / tot_i).mean().item(), "issuer_clocks_mean": tot_i.mean().item()}
            mma_clk = (3 * 8 + (4 * 0.5 if name == "bwd" else 0)) * 4 * terms * 64.0 * (tiles / float(sm_count()))
            st["mma_floor_share"] = mma_clk / st["issuer_clocks_mean"]
            results["%s_terms%d" % (name, terms)] = dict(ms=ms, us_per_tile_round=1e3 * ms / (tiles / 148.0),
                                                         tflops_alg=flop / ms / 1e9, stalls=st)
            print("terms=%d %s  %.3f ms  %.2f us/tile-round  %.1f TFLOP/s(alg)  | issuer: weights %.0f%% A %.0f%% Dfree "
                  "%.0f%%  MMA floor %.0f%% | epilogue: accum %.0f%% Xfree %.0f%% | producer: slot %.0f%%" % (
                      terms, name, ms, 1e3 * ms / (tiles / 148.0), flop / ms / 1e9, 100 * st["issuer_wait_weights"],
                      100 * st["issuer_wait_A"], 100 * st["issuer_wait_D_drained"], 100 * st["mma_floor_share"],
                      100 * st["epilogue_wait_accum"], 100 * st["epilogue_wait_X_free"], 100 * st["producer_wait_slot"]),
                  flush=True)
    if out_json:
        with open(out_json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
