#!/usr/bin/env python
"""bench.py -- rays/s of RobIR's per-ray rendering hot path on synthetic inputs of BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c2|c3|c4|c5|c2e] [--impl reference|reference-cuda]

Default = BASELINE config 2 (the configuration the metric is quoted on): each step = one training iteration of the
reference's PBR stage (training/train_pbr.py:431-460): 1024 random pixels of one 800x800 view, M=128 light SGs, S=32
samples/lobe: camera rays -> octree trace -> SDF normals -> material / indirect nets -> fused visibility MLP -> SG
render -> loss -> backward -> Adam step.  The other configurations (parity-test cases in the contract, measured here
for the record under profiles/):

    c1  64x64 crop (4096 rays, one call), M=16, forward only; both tracers (octree = shipped default, sphere tracer with
        ray_tracer.n_steps = 32 = the path the "march steps" knob controls)
    c3  Vis stage step (training/train_visibility.py:286-324): forward('Illum') + trace_radiance(nsamp=512) on 256 primary
        rays + IllumLoss + both backwards + both Adam steps; primary and secondary rays/s
    c4  PBR + CESR step (training/train_cesr.py:465-559, explore phase, S=8, lin_diff) -- ray-sharded over the ranks
    c5  the PBR step on a DTU-sized view (1600x1200, f=2892), ray-sharded over the ranks + gradient all-reduce

One process per GPU (torchrun for N>1), weak scaling: every rank renders its own batch and the gradient of the trained
parameters is all-reduced over NCCL each step.  Rank 0 prints ONE JSON line.

``--impl reference`` times the reference's OWN implementation of the path on the host cores: the unmodified RobIR code
staged under oracle/_ref (oracle/stage_ref.py; kind "reference"), or -- when no staged copy exists -- the oracle port
(kind "port"), on a pinned ray sample of the same workload.  ``--impl reference-cuda`` runs the same unmodified code
eagerly on the B200 (what RobIR's users run today); the default arm also reports it as ``cuda_baseline``.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from robir_b200 import synthetic  # noqa: E402

METRIC = "rays/sec (fwd+bwd) PBR stage, hotdog 800x800"
N_RAYS, M_LOBES, SDF_RADIUS, SEED = 1024, 128, 0.87, 0
FLOP_PER_QUERY = 458752.0  # 2 * (126*256 + 3*256^2 + 256*2)
REF_SAMPLE_RAYS = 256      # --impl reference: rays per step of the CPU arm (pinned: comparable across runs and N)

CAMERAS = {"hotdog": dict(H=800, W=800, focal=1111.1), "dtu": dict(H=1200, W=1600, focal=2892.0)}


def workload_config(config="c2", extra=None):
    cam = CAMERAS["dtu" if config == "c5" else "hotdog"]
    names = {"c2": "hotdog-synthetic 800x800 PBR stage: 1024 random pixels/step, M=128 light SGs, S=32, 24 indirect SGs, "
                   "octree tracer, geometric-init NeuS SDF (stage-2 radius ~0.6), fwd+loss+bwd+Adam",
             "c5": "dtu-synthetic 1600x1200 (f=2892) PBR stage: 1024 random pixels/step/GPU, M=128 light SGs, S=32, 24 "
                   "indirect SGs, octree tracer, geometric-init NeuS SDF (stage-2 radius ~0.6), fwd+loss+bwd+Adam"}
    cfg = {"workload": names.get(config, config), "config": config, "rays_per_step_per_gpu": N_RAYS,
           "num_lgt_sgs": M_LOBES, "image": "%dx%d" % (cam["W"], cam["H"]), "tracer": "octree",
           "l2": "per-step working set (ReLU masks + pair lists ~0.6 GB) exceeds the 126 MB L2; no explicit flush"}
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled with NVML from a background thread during the timed region."""

    def __init__(self, index, period=0.02):
        self.index, self.period, self.rows, self.stop_flag, self.thread = index, period, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[0].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def loop():
                while not self.stop_flag:
                    try:
                        self.rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                          pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
                    except Exception:
                        pass
                    time.sleep(self.period)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
        return self

    def stop(self):
        self.stop_flag = True
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.thread.join()
        import pynvml
        names = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for n, bit in names.items() if any(r[1] & bit for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def latest_ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed
    `ncu --set full` summary under profiles/ (a profiler number can only come from a profiler run; the file is named in
    the record)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_vis_tc.json")), reverse=True):
        try:
            d = json.load(open(path))
            if key in d and "dram_bytes" in d[key]:
                return d[key]["dram_bytes"], os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


def host_batch(step, cam, n=N_RAYS):
    pix = synthetic.training_pixels(step, n=n, H=cam["H"], W=cam["W"])
    uv = torch.stack([(pix % cam["W"]).float(), (pix // cam["W"]).float()], -1)[None]
    return uv, torch.ones(1, n, dtype=torch.bool), torch.full((1, n, 3), 0.5)


# ----------------------------------------------------------------------------------------------------------------------
# reference arms (CPU: --impl reference; eager CUDA: --impl reference-cuda / cuda_baseline)
# ----------------------------------------------------------------------------------------------------------------------
def _oracle_paths():
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)


def reference_available():
    _oracle_paths()
    import ref_shim
    return ref_shim.available()


def oracle_setup(sd, tree_arrays=None):
    _oracle_paths()
    import pipeline as P
    import robir_oracle as O
    import tracers as T
    if tree_arrays is None:
        cache = "/tmp/robir_oracle_octree_r%.2f_s%d.pt" % (SDF_RADIUS, SEED)
        if os.path.exists(cache):
            tree = torch.load(cache, weights_only=False)
        else:
            tree = T.OctreeOracle(lambda x: O.implicit_forward(sd, x)[:, 0],
                                  lambda x: O.implicit_gradient(sd, x)[:, 0, :])
            try:
                torch.save(tree, cache)
            except Exception:
                pass
    else:
        tree = T.OctreeOracle.__new__(T.OctreeOracle)
        for k, v in tree_arrays.items():
            setattr(tree, k, v)
        tree.max_iter = -1
    return O, P, tree


def oracle_step(O, P, tree, sd, step, n_rays, gen, cam):
    """One reference-equivalent training iteration of the oracle PORT on the CPU: forward + loss + backward."""
    pix = synthetic.training_pixels(step, n=N_RAYS, H=cam["H"], W=cam["W"])[:n_rays]
    inp = synthetic.camera_inputs(pix, **cam)
    train = [k for k in sd if k.startswith("envmap_material_network.") or k.startswith("gamma.")]
    for k in train:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    inp["hdr_shift"] = O.hdr_shift_as_input(sd).expand(n_rays, 1)
    gt = torch.full((1, n_rays, 3), 0.5)
    out = P.idr_forward(sd, inp, lambda c, m, d: tree.trace(c, d), lambda n: P.draw_rnd(n, M_LOBES, gen))
    loss, _ = O.pbr_loss(sd, out, gt)
    loss.backward()
    return int(out["network_object_mask"].sum())


def reference_pbr(device, sd):
    """The unmodified reference's PBR runner objects (oracle/ref_runner.py), octree built by the reference.  The
    reference prints progress to stdout; this script's stdout carries exactly one JSON line, so it goes to stderr."""
    import contextlib
    _oracle_paths()
    import ref_runner
    with contextlib.redirect_stdout(sys.stderr):
        R = ref_runner.ReferencePBR(sd, M_LOBES, device=device)
        R.generate()
    return R


def run_reference(args):
    """--impl reference: rank 0 only, host cores only (all of them)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.config not in ("c2", "c5"):
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm is implemented for the PBR step "
                                                              "(configs c2, c5); see --impl reference-cuda"}))
        return
    cam = CAMERAS["dtu" if args.config == "c5" else "hotdog"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=M_LOBES, sdf_radius=SDF_RADIUS)
    n_rays = REF_SAMPLE_RAYS
    if reference_available():
        kind = "reference"
        R = reference_pbr("cpu", sd)
        torch.manual_seed(1234)

        def step(s):
            pix = synthetic.training_pixels(s, n=N_RAYS, H=cam["H"], W=cam["W"])[:n_rays]
            inp = synthetic.camera_inputs(pix, **cam)
            inp.pop("hdr_shift")
            R.step(inp, {"rgb": torch.full((1, n_rays, 3), 0.5)})
    else:
        kind = "port"
        O, P, tree = oracle_setup(sd)
        gen = torch.Generator().manual_seed(1234)

        def step(s):
            oracle_step(O, P, tree, sd, s, n_rays, gen, cam)
    for s in range(args.warmup):
        step(s)
    t0 = time.time()
    for s in range(args.steps):
        step(100 + s)
    dt = time.time() - t0
    value = n_rays * args.steps / dt
    sample = ("%d of the %d rays of each step (the first %d of the step's random pixels of the %dx%d view), %d steps; %s"
              % (n_rays, N_RAYS, n_rays, cam["W"], cam["H"], args.steps,
                 "unmodified reference (oracle/_ref) incl. its Adam step" if kind == "reference"
                 else "oracle port, no optimizer step"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, {"rays_per_step_sample": n_rays}),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def reference_cuda_measure(sd, cam, steps, warmup):
    """The unmodified reference eagerly on cuda:0: full 1024-ray training iterations incl. H2D of the batch (its loop
    copies the inputs every iteration, train_pbr.py:438-439) and its own .item() syncs.  Returns a record."""
    R = reference_pbr("cuda", sd)
    torch.manual_seed(1234)
    batches = [host_batch(s, cam) for s in range(steps + warmup)]

    def step(s):
        uv, om, gt = batches[s]
        inp = {"uv": uv, "object_mask": om, "pose": synthetic.camera_pose(),
               "intrinsics": synthetic.camera_intrinsics(**cam)}
        return R.step(inp, {"rgb": gt})
    for s in range(warmup):
        step(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    hits = 0
    for s in range(warmup, warmup + steps):
        out, _ = step(s)
        hits += int(out["network_object_mask"].sum())
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    return {"value": N_RAYS * steps / dt, "unit": "rays/s", "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "kind": "reference-eager-cuda", "wall_ms_per_step": 1e3 * (time.time() - t0) / steps,
            "hit_fraction": hits / float(N_RAYS * steps),
            "what": "unmodified reference (oracle/_ref: IDRNetwork.forward + PBRTrainRunner.get_sg_render/pbr_step + "
                    "InvLoss + torch Adam), eager PyTorch on this GPU, its own octree, host batches, 1024 rays/step"}


def run_reference_cuda(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not reference_available() or not torch.cuda.is_available():
        print(json.dumps({"impl": "reference-cuda", "unavailable": "needs the staged reference (oracle/_ref) and a GPU"}))
        return
    cam = CAMERAS["dtu" if args.config == "c5" else "hotdog"]
    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=M_LOBES, sdf_radius=SDF_RADIUS)
    rec = reference_cuda_measure(sd, cam, args.steps, max(args.warmup, 2))
    print(json.dumps({
        "impl": "reference-cuda", "metric": METRIC, "value": rec["value"], "unit": "rays/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config), "cuda_baseline": rec,
        "e2e": {"value": rec["value"], "unit": "rays/s", "h2d_bytes_per_step": 21504, "d2h_bytes_per_step": 4}}))


# ----------------------------------------------------------------------------------------------------------------------
# configs c2 / c5: the PBR training step
# ----------------------------------------------------------------------------------------------------------------------
def run_pbr(args):
    import robir_b200
    from robir_b200 import _lib, dist as rdist, ops, rng
    from robir_b200.loss import InvLoss, pbr_step_loss
    cam = CAMERAS["dtu" if args.config == "c5" else "hotdog"]
    rank, world, local = rdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.engine:
        ops.ENGINE["vis"] = args.engine
    rng.set_mode("device")   # randoms are drawn on the GPU (the reference draws on the CPU and copies; same math)

    sd = synthetic.synthetic_state_dict(SEED, num_lgt_sgs=M_LOBES, sdf_radius=SDF_RADIUS)
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=M_LOBES)))
    model.load_state_dict(sd, strict=True)
    model.to(dev).train()
    model.generate()                                   # octree build: excluded from the metric
    loss_fn = InvLoss()
    params = list(model.gamma.parameters()) + list(model.envmap_material_network.parameters())
    opt = torch.optim.Adam(params, lr=5e-4, capturable=True,
                           fused=bool(int(os.environ.get("ROBIR_FUSED_ADAM", "1"))))   # train_pbr.py:104-105
    reducer = rdist.GradAllReducer(params)
    pose, K = synthetic.camera_pose().to(dev), synthetic.camera_intrinsics(**cam).to(dev)

    def pinned_batch(step):
        return tuple(t.pin_memory() for t in host_batch(step * world + rank, cam))

    def train_step(uv, om, gt):
        inp = {"uv": uv, "object_mask": om, "pose": pose, "intrinsics": K,
               "hdr_shift": model.gamma.hdr_shift.as_input().expand(N_RAYS, 1)}
        out = model(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss, _ = pbr_step_loss(model, loss_fn, out, {"rgb": gt})
        opt.zero_grad(set_to_none=True)
        loss.backward()
        reducer()
        opt.step()
        ops.invalidate_packed_weights()     # fused Adam updates in place without bumping tensor versions
        return loss, out["network_object_mask"]

    graphed = None
    if args.mode == "eager-static":
        model.static_shapes = True
        loss_fn.static_shapes = True
    if args.mode == "graph":
        from robir_b200.graph import GraphedPBRStep
        graphed = GraphedPBRStep(model, loss_fn, opt, N_RAYS, pose, K, reducer=reducer if world > 1 else None,
                                 split_reduce=args.split_reduce, pipeline_trace=args.pipeline_trace)

        def train_step(uv, om, gt, nxt=None):   # noqa: F811  (replay of the captured step)
            if args.pipeline_trace:             # nxt = (uv, object_mask) of the batch of the NEXT step
                return graphed(uv, om, gt, nxt[0], nxt[1]), None
            return graphed(uv, om, gt), None

    total = args.warmup + args.steps
    batches = [pinned_batch(s) for s in range(total)]
    dev_batches = [tuple(t.to(dev) for t in b) for b in batches]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, first, last):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(first, last):
            fn(s)
        e1.record()
        barrier()
        return rdist.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)

    # ---- (1) device-resident inputs: the headline `value`
    hits = []
    ops.Stats.reset()
    _lib.launch_count = 0

    pipelined = graphed is not None and args.pipeline_trace

    def step_resident(s):
        if pipelined:
            loss, m = train_step(*dev_batches[s % total], nxt=dev_batches[(s + 1) % total][:2])
        else:
            loss, m = train_step(*dev_batches[s % total])
        hits.append(graphed.hits.clone() if graphed is not None else m.sum())

    for s in range(args.warmup):
        step_resident(s)
    hits.clear()
    ops.Stats.reset()
    launches_before = _lib.launch_count
    clocks = ClockSampler(local).start()
    t_res = timed(step_resident, args.warmup, total)
    clk = clocks.stop()
    launches = _lib.launch_count - launches_before
    if graphed is not None:
        launches = graphed.launches_per_step * args.steps
    pairs_total = ops.Stats.total()
    n_hits = int(torch.stack(hits).sum().item())
    n_hits_all = rdist.sum_over_ranks(n_hits, dev)
    value = rdist.sum_over_ranks(N_RAYS * args.steps, dev) / t_res

    # ---- (2) end to end through the public API with host buffers: H2D of the batch + D2H of the loss inside the region
    h2d = sum(t.numel() * t.element_size() for t in batches[0])
    losses = []

    # Every step copies its loss to pinned host memory (D2H inside the timed region) and the host reads it one step later,
    # while the next step runs -- the logging pattern of a training loop that does not stall the device on .item() (the
    # reference prints its loss every 50 iterations, train_pbr.py:451-457; here every step's loss is read).
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]
    pending = []

    def read_pending():
        while pending:
            j = pending.pop(0)
            loss_done[j].synchronize()
            losses.append(float(loss_host[j]))

    def prime(s):                # pipelined mode: the trace buffers must hold batch s before a loop starts at s
        if pipelined:
            graphed.prime(*dev_batches[s % total][:2])

    def step_e2e(s):
        if pipelined:
            # this step's ground truth and the NEXT step's pixels / mask travel now (the walk of batch s + 1 runs under
            # this step's backward); the current batch's pixels were uploaded one step earlier
            gt = batches[s][2].to(dev, non_blocking=True)
            uv_n, om_n = (t.to(dev, non_blocking=True) for t in batches[(s + 1) % total][:2])
            loss, _ = train_step(dev_batches[s][0], dev_batches[s][1], gt, nxt=(uv_n, om_n))
        else:
            uv, om, gt = (t.to(dev, non_blocking=True) for t in batches[s])
            loss, _ = train_step(uv, om, gt)
        j = s & 1
        loss_host[j].copy_(loss.detach().reshape(()), non_blocking=True)     # device -> host copy of the step's result
        loss_done[j].record()
        read_pending()                                                      # host read of the previous step's loss
        pending.append(j)

    prime(0)
    for s in range(min(args.warmup, 3)):
        step_e2e(s)
    read_pending()
    losses.clear()
    prime(args.warmup)
    t_e2e = timed(step_e2e, args.warmup, total)
    read_pending()
    e2e = rdist.sum_over_ranks(N_RAYS * args.steps, dev) / t_e2e

    # ---- (2b) sustained: >= 5 s of back-to-back replays (thousands of steps), clocks sampled -- the timed region above
    # is ~70 ms at boost clocks; this is what a real training run sees under the power cap
    sustained = None
    if args.sustain > 0:
        n_sus = int(max(50, args.sustain / max(t_res / args.steps, 1e-4)))
        clocks = ClockSampler(local, period=0.05).start()
        prime(0)
        t_sus = timed(step_resident, 0, n_sus)
        clk_sus = clocks.stop()
        sustained = {"value": rdist.sum_over_ranks(N_RAYS * n_sus, dev) / t_sus, "unit": "rays/s", "steps": n_sus,
                     "seconds": t_sus, "ms_per_step": 1e3 * t_sus / n_sus, "clocks": clk_sus}

    # ---- (3) dominant kernel, timed live with CUDA events on the launching stream (eager replays of the same step)
    if graphed is not None:
        def train_step(uv, om, gt):   # noqa: F811
            loss = graphed._fwd_bwd()
            opt.step()
            ops.invalidate_packed_weights()
            return loss, None
    ops.PROFILE = []
    ops.Stats.reset()
    prof_steps = min(args.steps, 5)
    for s in range(args.warmup, args.warmup + prof_steps):
        train_step(*dev_batches[s])
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    # the dominant launch = the forward over the per-lobe diffuse pair list (one per step; the BRDF-lobe lists are a
    # second, ~100x smaller launch of the same kernel and are left out of the roofline figure)
    pairs_prof = ops.Stats.diffuse()
    big = [(n, a, b, t) for n, a, b, t in prof if t > 2048]
    t_fwd = sum(a.elapsed_time(b) for n, a, b, _ in big if n == "vis_mlp_fwd") * 1e-3
    t_bwd = sum(a.elapsed_time(b) for n, a, b, _ in big if n == "vis_mlp_bwd") * 1e-3
    n_fwd = sum(1 for n, *_ in big if n == "vis_mlp_fwd")
    peaks, peak_kind = measured_peaks()
    # the kernel is timed alone by CUDA events inside a ~20 ms eager region at boost clocks: the BURST figure is the
    # denominator (the sustained one is quoted beside it)
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    achieved = FLOP_PER_QUERY * pairs_prof / max(t_fwd, 1e-9) / 1e12
    terms = ops.vis_engine_terms() if hasattr(ops, "vis_engine_terms") else 3
    traffic, traffic_src = latest_ncu_traffic("fwd_diffuse")
    roofline = {"bound": "tensor", "kernel": "vis_tc_kernel<0> (visibility MLP forward, diffuse pair list)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": "%s bf16 burst (MEASURED_PEAKS.json: bf16_tflops); sustained %.1f" %
                               (peak_kind, peaks.get("bf16_tflops_sustained", 0.0)),
                "frac_of_sustained": achieved / peaks.get("bf16_tflops_sustained", peak_tf),
                "algorithmic_flop_per_launch": FLOP_PER_QUERY * pairs_prof / max(n_fwd, 1),
                "avg_launch_ms": 1e3 * t_fwd / max(n_fwd, 1), "launches_timed": n_fwd,
                "executed_tflops": achieved * (terms * 3 * 65536 + 512) / (126 * 256 + 3 * 65536 + 512.0),
                "note": "engine '%s': %d MMA term(s) per logical product (3 = fp32-parity split, 1 = single-pass fast "
                        "mode); executed tensor FLOP/s = executed_tflops" % (ops.ENGINE["vis"], terms),
                "bwd_kernel": {"achieved": FLOP_PER_QUERY * pairs_prof / max(t_bwd, 1e-9) / 1e12,
                               "frac": FLOP_PER_QUERY * pairs_prof / max(t_bwd, 1e-9) / 1e12 / peak_tf,
                               "share_of_step": t_bwd / prof_steps / (t_res / args.steps)},
                "share_of_step": t_fwd / prof_steps / (t_res / args.steps)}

    # ---- (3b) the octree walk (32 B per node visit), timed alone with CUDA events on the bench batch
    octree = None
    try:
        uv0 = dev_batches[args.warmup][0]
        rd, cl = ops.camera_rays(uv0, pose, K)
        tracer = model.ray_tracer
        tr = lambda: tracer(sdf=None, cam_loc=cl, object_mask=None, ray_directions=rd)
        for _ in range(3):
            tr()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            tr()
        e1.record()
        torch.cuda.synchronize()
        cnt = tracer.last_counters.cpu().tolist()
        visits, samples, iters = cnt[-8], cnt[-7], cnt[-6]        # [kMaxIter + 0 / 1 / 2] of kMaxIter + 8 counters
        t_oct = e0.elapsed_time(e1) * 1e-3 / 10
        hbm = peaks.get("hbm_gbs", 6650.0)
        octree = {"kernel": "octree_cast_kernel", "bound": "latency", "ms": 1e3 * t_oct, "rays": int(rd.shape[1]),
                  "lockstep_iterations": iters, "node_visits": visits, "micro_march_samples": samples,
                  "algorithmic_bytes": 32 * visits, "achieved_GBps": 32 * visits / t_oct / 1e9,
                  "frac_of_hbm_peak": 32 * visits / t_oct / 1e9 / hbm,
                  "us_per_iteration": 1e6 * t_oct / max(iters, 1),
                  "note": "one warp per ray, lock-step iterations (the reference's batch-level sample count depends on the "
                          "number of live rays every iteration): each iteration is a grid barrier plus a chain of "
                          "dependent 32-byte node fetches (L2-resident tree), so the bound is latency per iteration, not "
                          "bytes"}
    except Exception as exc:            # diagnostic only
        octree = {"error": str(exc)[:200]}

    # ---- (4) baselines on rank 0 at N=1: the reference's CPU path on a bounded sample, and the reference eagerly on
    # this GPU (what RobIR's users run today)
    cpu = cuda_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # a separate process: the reference needs its import shim in CPU mode (every .cuda() rewritten), which cannot
        # coexist with the CUDA-mode shim of the cuda_baseline below nor with this process's thread settings
        import subprocess
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config,
                                "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=900,
                               env=dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="0", WORLD_SIZE="1"))
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            cpu = json.loads(line)["cpu_baseline"]
        except Exception as exc:                                             # noqa: BLE001
            cpu = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    if rank == 0 and world == 1 and not args.no_cuda_baseline and reference_available():
        try:
            # free our graph's memory pool first? not needed: 180 GB; the reference keeps ~10 GB of activations
            cuda_ref = reference_cuda_measure(sd, cam, steps=5, warmup=2)
        except Exception as exc:                                             # noqa: BLE001
            cuda_ref = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        hit_frac = n_hits_all / float(N_RAYS * args.steps * world)
        print(json.dumps({
            "metric": METRIC if args.config == "c2" else "rays/sec (fwd+bwd) PBR stage, dtu 1600x1200",
            "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, {
                "parallelism": "rays x%d (+ NCCL grad all-reduce%s)" % (
                    world, "" if world == 1 else (", eager between two graphs" if args.split_reduce else
                                                  ", captured inside the step graph")), "hit_fraction": hit_frac,
                "vis_queries_per_step": pairs_total / float(args.steps), "vis_engine": ops.ENGINE["vis"],
                "rng": "device", "mode": args.mode,
                "pipeline_trace": bool(pipelined) and "octree walk of batch i+1 runs inside step i's graph, under its "
                                                      "loss / backward (frozen SDF: nothing trained feeds the tracer)"}),
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": 1e3 * t_e2e / args.steps,
                    "note": "per step: H2D of uv / mask / rgb from pinned memory, the step, D2H of its loss into pinned "
                            "memory; the host reads each loss one step later (no device stall on .item())"},
            "hit_rays_per_s": value * hit_frac, "sustained": sustained, "octree": octree,
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "cuda_baseline": cuda_ref, "final_loss": losses[-1] if losses else None}))
    if world > 1:
        torch.distributed.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="robir_b200")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "c2e"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-baseline", action="store_true")
    ap.add_argument("--sustain", type=float, default=5.0, help="seconds of back-to-back replays for the sustained record")
    ap.add_argument("--engine", default=None, help="visibility-MLP engine: tc (default, fp32 parity) | tc1 (single-pass "
                                                   "fast mode) | ffma")
    ap.add_argument("--capture-reduce", dest="split_reduce", action="store_false",
                    help="multi-GPU: capture the gradient all-reduce inside the step graph instead of issuing it eagerly "
                         "between the forward/backward graph and the optimizer graph (same speed, see graph.py)")
    ap.add_argument("--no-pipeline-trace", dest="pipeline_trace", action="store_false",
                    help="graph mode: trace every batch at the start of its own step instead of walking the NEXT batch "
                         "through the octree under the current step's loss / backward (GraphedPBRStep(pipeline_trace=True), "
                         "the default here: same numbers, the latency-bound walk leaves the critical path)")
    ap.add_argument("--mode", default="graph", help="graph: whole step as one CUDA graph (default) | eager | eager-static")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-cuda":
        return run_reference_cuda(args)
    args.warmup = max(args.warmup, 3)
    if args.config in ("c2", "c5"):
        return run_pbr(args)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs
    return getattr(bench_configs, "run_" + args.config)(args)


if __name__ == "__main__":
    main()
