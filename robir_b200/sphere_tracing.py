"""IDR sphere tracer with the reference module interface (model/ray_tracing.py:6-100):
``RayTracing(**conf)(sdf=..., cam_loc=..., object_mask=..., ray_directions=...) -> (points, mask, dists)``.

The reference evaluates ``sdf`` (a Python callable) ~45 times per call on boolean-masked sub-batches; here the whole
march runs in persistent CUDA kernels that evaluate the SDF network inline (csrc/sphere_trace.cu), so the ``sdf``
argument is only used to find the network whose weights the kernels read: a bound method of
``networks.ImplicitNetworkMy`` (``model.implicit_network.sdf``), or any callable after ``tracer.bind(implicit_network)``
(what ``robir_b200.install`` does on a reference model)."""
import torch
import torch.nn as nn

from . import ops, rng
from ._lib import RobirError


class RayTracing(nn.Module):
    def __init__(self, object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=1,
                 sphere_tracing_iters=10, n_steps=100, n_rootfind_steps=8):
        super().__init__()
        self.object_bounding_sphere = object_bounding_sphere
        self.sdf_threshold = sdf_threshold
        self.sphere_tracing_iters = sphere_tracing_iters
        self.line_step_iters = line_step_iters
        self.line_search_step = line_search_step
        self.n_steps = n_steps
        self.n_secant_steps = n_rootfind_steps
        self._weights = None
        self.last_counters = None

    def bind(self, implicit_network):
        """Attach the SDF network whose folded weights the march kernels read (an ImplicitNetworkMy, ours or the
        reference's: anything with .neus_model.sdf_network holding lin0..lin8 weight_g / weight_v / bias)."""
        w = getattr(implicit_network, "_w", None)
        self._weights = w if isinstance(w, ops.SdfWeights) else ops.SdfWeights(implicit_network.neus_model.sdf_network)
        return self

    def _weights_for(self, sdf):
        owner = getattr(sdf, "__self__", None)
        w = getattr(owner, "_w", None)
        if isinstance(w, ops.SdfWeights):
            return w
        if self._weights is None:
            raise RobirError("RayTracing: the sdf callable is not bound to a robir_b200 ImplicitNetworkMy; call "
                             "tracer.bind(model.implicit_network) first (there is no Python-callable fallback)")
        return self._weights

    def forward(self, sdf=None, cam_loc=None, object_mask=None, ray_directions=None):
        uni = None
        if self.training:
            # minimal_sdf_points draws n_steps CPU-uniform offsets per call (model/ray_tracing.py:305)
            uni = rng.uniform((self.n_steps,), ray_directions.device)
        out = ops.sphere_trace(self._weights_for(sdf), cam_loc, ray_directions, object_mask,
                               radius=self.object_bounding_sphere, sdf_threshold=self.sdf_threshold,
                               line_search_step=self.line_search_step, line_step_iters=self.line_step_iters,
                               sphere_tracing_iters=self.sphere_tracing_iters, n_steps=self.n_steps,
                               n_secant_steps=self.n_secant_steps, training=self.training, uniform_steps=uni,
                               return_stats=True)
        self.last_counters = out[3]
        return out[0], out[1], out[2]
