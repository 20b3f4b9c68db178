"""install(): the seams of a LIVE reference IDRNetwork are re-bound (build container only; no compute without a GPU)."""
import sys

import pytest
import torch

import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def test_install_rebinds_reference_seams(synth_sd16):
    import robir_b200
    from robir_b200 import synthetic, tracing
    from robir_b200 import sg_render as ours
    model = ref_shim.build_reference_model(synthetic.neus_checkpoint_from(synth_sd16), num_lgt_sgs=16)
    model.load_state_dict(synth_sd16, strict=True)
    keys_before = list(model.state_dict().keys())
    from robir_b200 import integration
    robir_b200.install(model)
    try:
        _check(model, keys_before, synth_sd16)
    finally:
        integration.uninstall_modules()
    assert sys.modules["model.sg_render"].render_with_all_sg is not ours.render_with_all_sg


def _check(model, keys_before, synth_sd16):
    import robir_b200
    from robir_b200 import tracing
    from robir_b200 import sg_render as ours
    assert isinstance(model.ray_tracer, tracing.OctreeTracing) and model.ray_tracer.max_iter == -1
    assert isinstance(model.octree_ray_tracer, tracing.OctreeTracing) and model.octree_ray_tracer.max_iter == 32
    assert sys.modules["model.sg_render"].render_with_all_sg is ours.render_with_all_sg
    assert sys.modules["model.implicit_differentiable_renderer"].render_with_all_sg is ours.render_with_all_sg
    assert list(model.state_dict().keys()) == keys_before          # checkpoints stay interchangeable
    # the re-bound ops refuse to run on CPU tensors instead of falling back
    with pytest.raises(robir_b200.RobirError):
        model.implicit_network(torch.zeros(4, 3))
    # our stand-alone model accepts the reference's state dict and vice versa
    mine = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=16)))
    mine.load_state_dict(model.state_dict(), strict=True)
    model.load_state_dict(mine.state_dict(), strict=True)
