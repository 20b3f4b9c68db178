"""Whole-step CUDA graph for the PBR stage: camera rays -> octree trace -> nets -> fused visibility MLP -> SG render ->
loss -> backward (-> Adam) captured once and replayed, so the ~1000 small launches and all Python overhead of a step
collapse into one graph launch.  Needs ``model.static_shapes = True`` (no data-dependent shapes / host syncs) and
device-side random numbers."""
import torch

from . import ops, rng
from .loss import pbr_step_loss


class GraphedPBRStep:
    """step(uv [1,N,2], object_mask [1,N] bool, rgb_gt [1,N,3]) -> loss (0-d device tensor, valid until the next call).

    With ``reducer`` (multi-GPU gradient all-reduce, ``dist.GradAllReducer``) the step is two graphs -- forward+backward
    and the optimizer update -- with the NCCL collective issued eagerly in between (``split_reduce=True``, default), or
    ONE graph with the collective captured between the backward and the optimizer update (``split_reduce=False``).
    Measured on 2 B200s (profiles/r2_v15_*): 3.80 vs 3.82 ms/step -- the ~0.1 ms over the single-GPU step is the
    collective's own serial device time (gather + 3.4 MB all-reduce + scatter after the last gradient), not host
    latency, so capturing it buys nothing; only bucketing it under the visibility backward would."""

    def __init__(self, model, loss_fn, optimizer, n_rays, pose, intrinsics, reducer=None, warmup=3,
                 record_randoms=False, split_reduce=True, hook=None, pipeline_trace=False):
        """record_randoms: keep references to the random tensors drawn inside the captured step (``self.random_tape``, in
        draw order); after a replay they hold the numbers that replay used (test hook: nothing in the graph changes).

        hook: a ``cesr.ClusteredAlbedoHook`` already bound to ``model.get_sg_render`` -- the step is then the CESR stage's
        (train_cesr.py:465-559,387-430: shadow_net / normal_net, supervise term, explore / project loss weights).  The
        hook's schedule (warm-up / explore / project, the normal switch at iteration 1000) is host-side control flow:
        one graph is captured per phase, lazily, when ``hook.phase_key()`` changes.

        pipeline_trace: software-pipeline the surface trace across steps.  The octree walk is latency-bound (65 lock-step
        iterations, 0.6 % of the HBM peak) and depends on nothing that is trained, so step i's graph walks batch i + 1 on a
        side branch under its own loss / backward and step i + 1 starts from the finished trace.  Call ``prime(uv, om)``
        once with the first batch, then ``step(uv_i, om_i, gt_i, uv_next, om_next)``; ``uv_i`` / ``om_i`` must be what the
        previous call passed as next (they are only used by ``prime``).  Same numbers as the plain step."""
        if rng._mode != "device":
            raise RuntimeError("GraphedPBRStep needs robir_b200.rng.set_mode('device')")
        model.static_shapes = True
        loss_fn.static_shapes = True
        dev = pose.device
        self.model, self.loss_fn, self.opt, self.reducer, self.hook = model, loss_fn, optimizer, reducer, hook
        self.uv = torch.zeros(1, n_rays, 2, device=dev)
        self.om = torch.ones(1, n_rays, dtype=torch.bool, device=dev)
        self.gt = torch.zeros(1, n_rays, 3, device=dev)
        self.pose, self.K, self.n = pose, intrinsics, n_rays
        self.hits = None
        self.pipeline_trace = pipeline_trace
        if pipeline_trace:
            self.uv_next, self.om_next = torch.zeros_like(self.uv), torch.ones_like(self.om)
            self.cur = [torch.zeros(1, n_rays, 3, device=dev), torch.zeros(1, 3, device=dev),
                        torch.zeros(n_rays, dtype=torch.bool, device=dev), torch.zeros(n_rays, device=dev)]
            self.nxt = [torch.zeros_like(t) for t in self.cur]
            self.trace_stream = torch.cuda.Stream()
            self._primed = False
        self.multi = reducer is not None and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size() > 1
        self.split = self.multi and split_reduce
        self.record_randoms = record_randoms
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                if self.multi:
                    reducer()
                self.opt.step()
                ops.invalidate_packed_weights()   # optimizers with fused multi-tensor kernels do not bump versions
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graphs = {}
        self._capture(self._phase())

    def _phase(self):
        return self.hook.phase_key() if self.hook is not None else None

    def _capture(self, key):
        import contextlib
        from . import _lib
        g1, g2 = torch.cuda.CUDAGraph(), None
        before = _lib.launch_count
        with (rng.record(on_device=True) if self.record_randoms else contextlib.nullcontext()) as tape:
            with torch.cuda.graph(g1):
                loss = self._fwd_bwd()
                if not self.split:
                    if self.multi:
                        self.reducer()             # NCCL all-reduce as a node of this graph
                    self.opt.step()
        launches = _lib.launch_count - before      # kernels of the C-ABI library inside one replay
        if self.split:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                self.opt.step()
        # the gradient tensors this capture allocated (its backward writes them, its optimizer update and the eager
        # all-reduce of the split mode read them): p.grad is pointed back at them whenever this graph is selected
        grads = [(p, p.grad) for g in self.opt.param_groups for p in g["params"]]
        self._graphs[key] = dict(g1=g1, g2=g2, loss=loss, hits=self.hits, launches=launches, grads=grads,
                                 tape=list(tape) if self.record_randoms else None)
        self._select(key)

    def _select(self, key):
        g = self._graphs[key]
        self.g1, self.g2, self.loss, self.hits = g["g1"], g["g2"], g["loss"], g["hits"]
        self.launches_per_step, self.random_tape = g["launches"], g["tape"]
        for p, gr in g["grads"]:
            p.grad = gr
        self._key = key

    def _fwd_bwd(self):
        m = self.model
        ops.invalidate_packed_weights()      # the optimizer ran since the last forward: trained weights are re-packed once
        inp = {"uv": self.uv, "object_mask": self.om, "pose": self.pose, "intrinsics": self.K,
               "hdr_shift": m.gamma.hdr_shift.as_input().expand(self.n, 1)}
        if self.pipeline_trace:
            inp["traced"] = tuple(self.cur)
        out = m(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        if self.pipeline_trace:
            # the next batch's walk starts when this step's forward is done and runs under the loss and the backward
            main = torch.cuda.current_stream()
            self.trace_stream.wait_stream(main)
            with torch.cuda.stream(self.trace_stream):
                for dst, src in zip(self.nxt, m.trace_rays(self.uv_next, self.pose, self.K, self.om_next)):
                    dst.copy_(src)
        if self.hook is not None:
            loss, _ = self.hook.pbr_step(self.loss_fn, out, {"rgb": self.gt})
        else:
            loss, _ = pbr_step_loss(m, self.loss_fn, out, {"rgb": self.gt})
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.hits = out["network_object_mask"].sum()
        if self.pipeline_trace:
            torch.cuda.current_stream().wait_stream(self.trace_stream)
            for dst, src in zip(self.cur, self.nxt):          # every consumer of the current trace is behind us
                dst.copy_(src)
        return loss.detach()

    def prime(self, uv, object_mask):
        """pipeline_trace: trace the first batch (eagerly) into the buffers the next replay starts from."""
        with torch.no_grad():
            for dst, src in zip(self.cur, self.model.trace_rays(uv, self.pose, self.K, object_mask)):
                dst.copy_(src)
        self._primed = True

    def __call__(self, uv, object_mask, rgb_gt, uv_next=None, om_next=None):
        if self.pipeline_trace:
            if uv_next is None or om_next is None:
                raise RuntimeError("GraphedPBRStep(pipeline_trace=True): pass the next batch's uv / object_mask")
            if not self._primed:
                self.prime(uv, object_mask)
            self.uv_next.copy_(uv_next, non_blocking=True)
            self.om_next.copy_(om_next, non_blocking=True)
        key = self._phase()
        if key != self._key:
            if key in self._graphs:
                self._select(key)
            else:
                self._capture(key)
        if not self.pipeline_trace:          # the pipelined graph starts from the finished trace: it never reads uv
            self.uv.copy_(uv, non_blocking=True)
        self.om.copy_(object_mask, non_blocking=True)
        self.gt.copy_(rgb_gt, non_blocking=True)
        self.g1.replay()
        if self.g2 is not None:
            self.reducer()
            self.g2.replay()
        # the replay's optimizer update changed the trained weights without touching tensor versions or the host-side
        # epoch: an eager forward after training (plots, evaluation) must not hit the packed copies of the step before
        ops.invalidate_packed_weights()
        return self.loss
