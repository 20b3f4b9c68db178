"""robir_b200 -- B200-native (sm_100a) implementation of RobIR's per-ray rendering hot path behind the reference's own
Python module surface.  See DESIGN.md / INTEGRATION.md.

    import robir_b200
    model = robir_b200.IDRNetwork(conf)          # same forward()/trace_radiance()/state_dict() as the reference
    robir_b200.install(reference_model)          # or: re-bind the hot path on a live reference IDRNetwork
"""
from . import rng, synthetic  # noqa: F401
from ._lib import LIB_PATH, RobirError  # noqa: F401


def __getattr__(name):
    # heavy modules are imported lazily so that `import robir_b200.synthetic` works without touching the extension
    if name in ("IDRNetwork", "pbr_get_sg_render"):
        from . import renderer
        return getattr(renderer, name)
    if name in ("render_with_all_sg", "get_diffuse_visibility", "get_specular_visibility"):
        from . import sg_render
        return getattr(sg_render, name)
    if name in ("OctreeTracing",):
        from . import tracing
        return getattr(tracing, name)
    if name == "install":
        from .integration import install
        return install
    if name == "invalidate_packed_weights":
        from .ops import invalidate_packed_weights
        return invalidate_packed_weights
    if name == "render_image":
        from .render import render_image
        return render_image
    if name in ("InvLoss", "pbr_step_loss"):
        from . import loss
        return getattr(loss, name)
    raise AttributeError(name)
