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
                 record_randoms=False, split_reduce=True):
        """record_randoms: keep references to the random tensors drawn inside the captured step (``self.random_tape``, in
        draw order); after a replay they hold the numbers that replay used (test hook: nothing in the graph changes)."""
        if rng._mode != "device":
            raise RuntimeError("GraphedPBRStep needs robir_b200.rng.set_mode('device')")
        model.static_shapes = True
        loss_fn.static_shapes = True
        dev = pose.device
        self.model, self.loss_fn, self.opt, self.reducer = model, loss_fn, optimizer, reducer
        self.uv = torch.zeros(1, n_rays, 2, device=dev)
        self.om = torch.ones(1, n_rays, dtype=torch.bool, device=dev)
        self.gt = torch.zeros(1, n_rays, 3, device=dev)
        self.pose, self.K, self.n = pose, intrinsics, n_rays
        self.hits = None
        multi = reducer is not None and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        split = multi and split_reduce
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                if multi:
                    reducer()
                self.opt.step()
                ops.invalidate_packed_weights()   # optimizers with fused multi-tensor kernels do not bump versions
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        self.g1 = torch.cuda.CUDAGraph()
        self.g2 = None
        before = _lib.launch_count
        import contextlib
        with (rng.record(on_device=True) if record_randoms else contextlib.nullcontext()) as tape:
            with torch.cuda.graph(self.g1):
                self.loss = self._fwd_bwd()
                if not split:
                    if multi:
                        reducer()                  # NCCL all-reduce as a node of this graph
                    self.opt.step()
        self.random_tape = list(tape) if record_randoms else None
        self.launches_per_step = _lib.launch_count - before    # kernels of the C-ABI library inside one replay
        if split:
            self.g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g2):
                self.opt.step()

    def _fwd_bwd(self):
        m = self.model
        ops.invalidate_packed_weights()      # the optimizer ran since the last forward: trained weights are re-packed once
        inp = {"uv": self.uv, "object_mask": self.om, "pose": self.pose, "intrinsics": self.K,
               "hdr_shift": m.gamma.hdr_shift.as_input().expand(self.n, 1)}
        out = m(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss, _ = pbr_step_loss(m, self.loss_fn, out, {"rgb": self.gt})
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.hits = out["network_object_mask"].sum()
        return loss.detach()

    def __call__(self, uv, object_mask, rgb_gt):
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
