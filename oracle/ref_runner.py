"""TEST / BENCH INFRASTRUCTURE ONLY -- the UNMODIFIED reference's PBR training iteration, driven the way
``PBRTrainRunner.run`` drives it (training/train_pbr.py:398-460), on the host cores (device="cpu") or eagerly on the
GPU (device="cuda"), from ``/root/reference`` or its staged copy ``oracle/_ref/`` (oracle/stage_ref.py).

Every line of compute here is the reference's own code: ``IDRNetwork.forward`` (model/implicit_differentiable_renderer.py
:290-479), ``PBRTrainRunner.get_sg_render`` / ``pbr_step`` (training/train_pbr.py:318-396), ``InvLoss`` (model/loss.py),
``OctreeTracing.generate`` (model/octree_tracing.py:31-41) and torch.optim.Adam over the runner's parameter set
(train_pbr.py:104-105).  This module only builds the objects the runner's ``__init__`` would build from a dataset /
checkpoint directory that does not exist here (hard part 8).
"""
import os
import sys

import torch

import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class ReferencePBR:
    def __init__(self, state_dict, num_lgt_sgs, device="cpu", lr=5e-4, optimizer=True):
        from robir_b200 import synthetic
        self.device = device
        model = ref_shim.build_reference_model(synthetic.neus_checkpoint_from(state_dict), num_lgt_sgs=num_lgt_sgs,
                                               device=device)
        model.load_state_dict(state_dict, strict=True)
        if device == "cuda":
            model.cuda()                                                    # train_pbr.py:92-93
        self.model = model
        self.runner = ref_shim.bind_pbr_runner(model)
        from model.loss import InvLoss
        self.runner.loss = InvLoss(1.0, 0.1, 100.0, 50.0, 1.0, 1.0, 1.0)    # confs_sg/hotdog.conf:47-58
        model.train()
        self.opt = None
        if optimizer:                                                       # train_pbr.py:104-105
            self.opt = torch.optim.Adam(list(model.gamma.parameters()) +
                                        list(model.envmap_material_network.parameters()), lr=lr)

    def generate(self, secondary=False):
        """train_pbr.py:403-407 (no tex_sampler: the octree spans [-1, 1]^3, octree_tracing.py:33-37).  The PBR step
        never casts secondary rays (trace_vis = False, :399), so the second, identical octree is only built on request."""
        m = self.model
        m.ray_tracer.generate(lambda x: m.implicit_network(x)[:, 0], None)
        if secondary and hasattr(m, "octree_ray_tracer"):
            m.octree_ray_tracer.generate(lambda x: m.implicit_network(x)[:, 0], None)

    def forward_loss(self, model_input, ground_truth):
        """train_pbr.py:438-445."""
        m = self.model
        inp = {k: (v.cuda() if self.device == "cuda" else v) for k, v in model_input.items()}
        inp["hdr_shift"] = m.gamma.hdr_shift.as_input().expand(inp["uv"].shape[1], 1)
        out = m(inp, trainstage="Material", fun_spec=False, lin_diff=False, train_spec=True)
        loss, _ = self.runner.pbr_step(out, ground_truth)
        return out, loss

    def step(self, model_input, ground_truth):
        """One training iteration: train_pbr.py:438-449."""
        out, loss = self.forward_loss(model_input, ground_truth)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return out, loss
