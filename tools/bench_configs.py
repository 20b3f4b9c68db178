"""bench.py --config c1 | c3 | c4: the BASELINE.json configurations besides the headline one (c2) and its DTU-sized
sibling (c5), which live in bench.py itself.  Same JSON contract (one line on rank 0, CUDA-event timing, max over ranks,
weak scaling over the ranks); these are measured for the record under profiles/, the driver's bench is c2.

    c1  64x64 centre crop (4096 rays, one call), M = 16 light SGs, PBR forward only (no_grad) -- octree tracer (shipped
        default; headline value) and the IDR sphere tracer with ray_tracer.n_steps = 32 (the "march steps" knob)
    c3  Vis stage iteration (training/train_visibility.py:286-324): forward('Illum') on 256 primary rays +
        trace_radiance(nsamp = 512) + IllumLoss + both backwards + both Adam steps; primary and secondary rays/s
    c2e full 800x800 image of config 2's model in evaluation mode (PBRTrainRunner.plot_to_disk, training/train_pbr.py:235-311):
        forward only, the reference's 1024-pixel chunks and 16 384-pixel chunks; a "step" is one whole image
    c4  PBR + CESR iteration (training/train_cesr.py:465-559, explore phase, S = 8, lin_diff), every rank its own batch,
        one CUDA graph per step (--mode eager: the dynamic-shape eager step)
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from robir_b200 import synthetic  # noqa: E402

SEED, SDF_RADIUS = 0, 0.87


def _setup(M, use_octree=True, n_steps=100):
    import robir_b200
    from robir_b200 import dist as rdist, rng
    rank, world, local = rdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rng.set_mode("device")
    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=M, sdf_radius=SDF_RADIUS)
    conf = dict(envmap_material_network=dict(num_lgt_sgs=M), use_octree=use_octree,
                ray_tracer=dict(n_steps=n_steps))
    model = robir_b200.IDRNetwork(conf)
    model.load_state_dict(sd, strict=True)
    model.to(dev).train()
    model.generate()
    return rank, world, local, dev, sd, model


def _timed(fn, warmup, steps, dev, world):
    from robir_b200 import dist as rdist
    for s in range(warmup):
        fn(s)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(warmup, warmup + steps):
        fn(s)
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    return rdist.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)


def _line(args, metric, value, unit, ms, world, cfg, extra):
    d = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
         "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
         "data": "synthetic", "config": cfg}
    d.update(extra)
    print(json.dumps(d))


# ----------------------------------------------------------------------------------------------------------------------
def run_c1(args):
    """64x64 crop, 16 SG lobes, PBR forward only."""
    import bench
    N, M = 4096, 16
    out = {}
    for tracer, use_octree in (("octree", True), ("sphere_tracer_n_steps_32", False)):
        rank, world, local, dev, sd, model = _setup(M, use_octree=use_octree, n_steps=32)
        model.eval() if False else None          # the reference times the training-mode forward
        yy, xx = torch.meshgrid(torch.arange(368, 432), torch.arange(368, 432), indexing="ij")
        pix = (yy * 800 + xx).reshape(-1)
        inp = {k: v.to(dev) for k, v in synthetic.camera_inputs(pix).items()}
        hits = []

        def step(s):
            with torch.no_grad():
                o = model(inp, trainstage="Material", train_spec=True)
            hits.append(o["network_object_mask"].sum())
        clocks = bench.ClockSampler(local).start()
        t = _timed(step, args.warmup, args.steps, dev, world)
        clk = clocks.stop()
        out[tracer] = dict(rays_per_s=N * args.steps * world / t, ms_per_call=1e3 * t / args.steps,
                           hit_fraction=float(torch.stack(hits[-args.steps:]).float().mean()) / N, clocks=clk)
    if rank == 0:
        cfg = {"workload": "hotdog-synthetic 64x64 centre crop (4096 rays, one call), M=16 light SGs, PBR forward only "
                           "(no_grad, training-mode networks), stage-2 SDF radius ~0.6", "config": "c1",
               "rays_per_call": N, "num_lgt_sgs": M, "tracer": "octree"}
        o = out["octree"]
        _line(args, "rays/sec (fwd) PBR stage, hotdog 64x64 crop, 16 SG lobes", o["rays_per_s"], "rays/s",
              o["ms_per_call"], world, cfg, {"clocks": o["clocks"], "tracers": out,
                                             "e2e": {"value": o["rays_per_s"], "unit": "rays/s",
                                                     "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                                     "note": "inputs resident; forward-only diagnostic config"}})


# ----------------------------------------------------------------------------------------------------------------------
def run_c3(args):
    """Vis stage iteration: 256 primary rays x 512 secondary rays."""
    import bench
    from robir_b200.loss import IllumLoss
    N, NS, M = 256, 512, 128
    rank, world, local, dev, sd, model = _setup(M)
    model.indirect_illum_network.train_weights = True
    illum_loss = IllumLoss()
    illum_opt = torch.optim.Adam(model.indirect_illum_network.parameters(), lr=5e-4)     # train_visibility.py:99-107
    vis_opt = torch.optim.Adam(model.visibility_network.parameters(), lr=5e-4)
    stats = {"sec": 0, "sec_hit": 0, "prim_hit": 0}
    # host batches prepared before the timed region (the runner's DataLoader does this work off the critical path)
    host = [synthetic.camera_inputs(synthetic.training_pixels(s * world + rank, n=N))
            for s in range(args.warmup + args.steps)]

    def step(s):
        inp = {k: v.to(dev, non_blocking=True) for k, v in host[s].items()}
        inp["hdr_shift"] = torch.rand(N, 1, device=dev)                                   # :297
        out = model(inp, trainstage="Illum")
        tr = model.trace_radiance(out, nsamp=NS)
        n_hit = int(out["network_object_mask"].sum())
        if n_hit:
            rad, vis = illum_loss(out, tr, 0.0)
            vis_opt.zero_grad()
            vis.backward(retain_graph=True)
            vis_opt.step()
            illum_opt.zero_grad()
            rad.backward()
            illum_opt.step()
        stats["prim_hit"] += n_hit
        stats["sec"] += n_hit * NS
        stats["sec_hit"] += int(tr["gt_vis"].sum())
    clocks = bench.ClockSampler(local).start()
    for k in stats:
        stats[k] = 0
    t = _timed(step, args.warmup, args.steps, dev, world)
    clk = clocks.stop()
    tot = args.warmup + args.steps
    if rank == 0:
        cfg = {"workload": "lego-synthetic Vis stage: forward('Illum') on 256 random pixels + trace_radiance(nsamp=512) + "
                           "IllumLoss + both backwards + both Adam steps (train_visibility.py:286-324), eager dynamic "
                           "shapes, stage-2 SDF radius ~0.6", "config": "c3", "primary_rays_per_step_per_gpu": N,
               "secondary_per_hit": NS, "primary_hit_fraction": stats["prim_hit"] / float(N * tot),
               "secondary_hit_fraction": stats["sec_hit"] / float(max(stats["sec"], 1))}
        sec_per_step = stats["sec"] / float(tot)
        _line(args, "primary rays/sec Vis stage (256 x 512 secondary), lego 800x800", N * args.steps * world / t,
              "rays/s", 1e3 * t / args.steps, world, cfg,
              {"secondary_rays_per_s": sec_per_step * args.steps * world / t, "clocks": clk,
               "e2e": {"value": N * args.steps * world / t, "unit": "rays/s", "h2d_bytes_per_step": N * 8 + N + 100,
                       "d2h_bytes_per_step": 8, "note": "the step builds its batch on the host and copies it (dynamic "
                                                        "shapes sync the host several times per step)"}})


# ----------------------------------------------------------------------------------------------------------------------
def run_c2e(args):
    """Whole 800x800 image, evaluation mode (train_pbr.py:235-311: model.eval(), is_training = False, split_input chunks,
    hdr2ldr of sg_rgb + indir_rgb), forward only.  One step = one image; the image is read back to the host (e2e)."""
    import bench
    H = W = 800
    M = 128
    rank, world, local, dev, sd, model = _setup(M)
    import robir_b200
    pose, K = synthetic.camera_pose().to(dev), synthetic.camera_intrinsics().to(dev)
    host_img = torch.empty(H * W, 3).pin_memory()
    bufs = None
    out = {}
    for chunk in (1024, 16384):
        state = {}

        def step(s, chunk=chunk):
            nonlocal bufs
            bufs = robir_b200.render_image(model, pose, K, H, W, chunk=chunk, out=bufs)      # the public entry point
            host_img.copy_(bufs["pred_rgb"], non_blocking=True)
            state["hits"] = bufs["network_object_mask"].sum()
        clocks = bench.ClockSampler(local).start()
        t = _timed(step, 1, args.steps, dev, world)
        clk = clocks.stop()
        torch.cuda.synchronize()
        out["chunk_%d" % chunk] = dict(rays_per_s=H * W * args.steps * world / t, s_per_image=t / args.steps,
                                       hit_fraction=float(state["hits"]) / (H * W), clocks=clk,
                                       finite=bool(torch.isfinite(host_img).all()))
    if rank == 0:
        best = out["chunk_16384"]
        cfg = {"workload": "hotdog-synthetic full 800x800 image, evaluation mode (plot_to_disk path: model.eval(), testing "
                           "visibility, hdr2ldr), M=128, S=32, forward only, eager dynamic shapes; headline = 16 384-pixel "
                           "chunks, the reference's 1024-pixel chunking beside it", "config": "c2e",
               "rays_per_image": H * W, "num_lgt_sgs": M, "chunks": out}
        _line(args, "rays/sec (fwd, eval) PBR stage full image, hotdog 800x800", best["rays_per_s"], "rays/s",
              1e3 * best["s_per_image"], world, cfg,
              {"clocks": best["clocks"], "e2e": {"value": best["rays_per_s"], "unit": "rays/s", "h2d_bytes_per_step": 0,
                                                 "d2h_bytes_per_step": H * W * 12,
                                                 "note": "pixel grid resident, rendered image copied to pinned host memory "
                                                         "inside the timed region"}})
    if world > 1:
        torch.distributed.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
def run_c4(args):
    """PBR + CESR iteration (explore phase), every rank its own 1024-ray batch.  --mode graph (default): fixed-capacity
    batch, the whole step (forward + CESR hook + loss + backward + Adam over 3 networks) as one CUDA graph;
    --mode eager: the dynamic-shape eager step (what round 1 measured)."""
    import bench
    from robir_b200 import cesr, dist as rdist, graph, ops
    from robir_b200.loss import InvLoss
    N, M = 1024, 128
    rank, world, local, dev, sd, model = _setup(M)
    model.static_shapes = False
    sh, nr = synthetic.cesr_state_dicts(SEED)
    shadow, normal = cesr.WnMLP(191, 2), cesr.WnMLP(63, 3)
    shadow.load_state_dict(sh)
    normal.load_state_dict(nr)
    hook = cesr.ClusteredAlbedoHook(model, shadow.to(dev), normal.to(dev), cur_iter=600)     # explore phase
    model.get_sg_render = hook.get_sg_render
    loss_fn = InvLoss()
    params = list(model.gamma.parameters()) + list(model.envmap_material_network.parameters()) + hook.parameters()
    graphed = args.mode == "graph"
    opt = torch.optim.Adam(params, lr=5e-4, capturable=graphed, fused=graphed)               # train_cesr.py:111-117
    reducer = rdist.GradAllReducer(params)
    pose, K = synthetic.camera_pose().to(dev), synthetic.camera_intrinsics().to(dev)
    hits = []
    host = [synthetic.training_pixels(s * world + rank, n=N) for s in range(args.warmup + args.steps)]
    om = torch.ones(1, N, dtype=torch.bool, device=dev)
    gt = torch.full((1, N, 3), 0.5, device=dev)
    if graphed:
        host_uv = [torch.stack([(p % 800).float(), (p // 800).float()], -1)[None].pin_memory() for p in host]
        pipe = getattr(args, "pipeline_trace", True)
        gstep = graph.GraphedPBRStep(model, loss_fn, opt, N, pose, K, reducer=reducer if world > 1 else None, hook=hook,
                                     pipeline_trace=pipe)
        n_b = len(host_uv)
        if pipe:
            gstep.prime(host_uv[0].to(dev), om)
        cur_uv = [host_uv[0].to(dev)]

        def step(s):
            if pipe:      # the NEXT batch's pixels travel now: its octree walk runs under this step's backward
                nxt = host_uv[(s + 1) % n_b].to(dev, non_blocking=True)
                gstep(cur_uv[0], om, gt, nxt, om)
                cur_uv[0] = nxt
            else:
                gstep(host_uv[s].to(dev, non_blocking=True), om, gt)
            hits.append(gstep.hits.clone())
    else:
        def step(s):
            pix = host[s].to(dev, non_blocking=True)
            uv = torch.stack([(pix % 800).float(), (pix // 800).float()], -1)[None]
            inp = {"uv": uv, "object_mask": om, "pose": pose,
                   "intrinsics": K, "hdr_shift": model.gamma.hdr_shift.as_input().expand(N, 1)}
            out = model(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
            loss, _ = hook.pbr_step(loss_fn, out, {"rgb": gt})
            opt.zero_grad(set_to_none=True)
            loss.backward()
            reducer()
            opt.step()
            hits.append(out["network_object_mask"].sum())
    clocks = bench.ClockSampler(local).start()
    t = _timed(step, args.warmup, args.steps, dev, world)
    clk = clocks.stop()
    prof_path = os.environ.get("ROBIR_C4_PROFILE")
    if prof_path and rank == 0:                 # per-kernel table of two more steps (diagnostic, after the timed region)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for s in range(2):
                step(args.warmup + args.steps - 1 - s)
            torch.cuda.synchronize()
        with open(prof_path, "w") as f:
            f.write(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=90))
    if rank == 0:
        hf = float(torch.stack(hits[-args.steps:]).float().mean()) / N
        cfg = {"workload": "truck-synthetic PBR + CESR step: 1024 random pixels/step/GPU, M=128, shadow_net (191->512x8->2 "
                           "on n_hit x 128 rows) + normal_net, explore phase (iteration 600), S=8, fwd+loss+bwd+Adam over 3 "
                           "networks", "config": "c4", "rays_per_step_per_gpu": N, "num_lgt_sgs": M,
               "mode": ("one CUDA graph per step (fixed-capacity batch)" + (", next batch's octree walk pipelined under the "
                        "backward" if getattr(args, "pipeline_trace", True) else "")) if graphed else "eager dynamic shapes",
               "launches_per_step": gstep.launches_per_step if graphed else None,
               "hit_fraction": hf, "wn_engine": ops.ENGINE["wn"],
               "parallelism": "rays x%d (+ NCCL grad all-reduce incl. shadow_net / normal_net)" % world}
        _line(args, "rays/sec (fwd+bwd) PBR+CESR stage, truck", N * args.steps * world / t, "rays/s",
              1e3 * t / args.steps, world, cfg,
              {"clocks": clk, "e2e": {"value": N * args.steps * world / t, "unit": "rays/s",
                                      "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": 0,
                                      "note": "pixel coordinates are drawn on the host and copied every step"}})
    if world > 1:
        torch.distributed.destroy_process_group()
