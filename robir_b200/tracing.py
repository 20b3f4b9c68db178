"""Surface tracers with the reference module interface (``tracer(sdf=..., cam_loc=..., object_mask=...,
ray_directions=...) -> (points, mask, dist)`` and ``generate(sdf_fn, tex_sampler=None)``).

OctreeTracing mirrors model/octree_tracing.py:8-60 over utils/octree.py (build :124-199, :377-409; cast :421-585):
the tree is built once with batched CUDA SDF evaluations and walked by one cooperative kernel per call.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import RobirError


def _child_boxes(boxes):
    c = torch.arange(8, device=boxes.device)
    ofs = torch.stack([(c // 4) % 2, (c // 2) % 2, c % 2], -1)
    mn = boxes[:, None, :3] + ofs * boxes[:, None, 3:] / 2
    sz = (boxes[:, None, 3:] / 2).expand(mn.shape)
    return torch.cat([mn, sz], -1)


def build_octree(sdf_and_grad, bounds, device, thr=0.5, cell_size=0.05, depth=4, chunk=1 << 18):
    """sdf_and_grad(x[k,3]) -> (sdf [k], grad [k,3]).  Returns ops.PackedOctree.  (utils/octree.py:124-199, 377-409)"""
    lo = torch.tensor(bounds[0], dtype=torch.float32, device=device)
    size = torch.tensor([bounds[1][i] - bounds[0][i] for i in range(3)], dtype=torch.float32, device=device)
    cells = (size / cell_size).ceil().long()
    size = cells * cell_size
    root = torch.cat([lo, size])
    axes = [torch.arange(int(c), device=device) for c in cells]
    anchor = torch.stack(torch.meshgrid(axes, indexing="ij"), -1).view(-1, 3)
    bmin = (anchor / cells) * root[3:] + root[:3]
    bmax = ((anchor + 1.0) / cells) * root[3:] + root[:3]
    boxes = torch.cat([bmin, bmax - bmin], -1)
    non_leaf = torch.zeros(boxes.shape[0], dtype=torch.long, device=device)
    links = -torch.ones(boxes.shape[0], 8, dtype=torch.long, device=device)
    grid = torch.arange(anchor.shape[0], device=device).view(*[int(c) for c in cells])

    def batched(x, want_grad):
        vals, grads = [], []
        for j in range(0, x.shape[0], chunk):
            s, g = sdf_and_grad(x[j:j + chunk].float().contiguous(), want_grad)
            vals.append(s)
            grads.append(g)
        return torch.cat(vals, 0), (torch.cat(grads, 0) if want_grad else None)

    start, end = 0, boxes.shape[0]
    for _ in range(depth):
        lvl = boxes[start:end]
        sz = lvl[:, 3:]
        sdf, _ = batched(lvl[:, :3] + sz * 0.5, False)
        split = sdf.abs() < sz.norm(dim=-1) * thr
        if not bool(split.any()):
            break
        k = split.nonzero()[:, 0]
        n = k.shape[0]
        non_leaf[start:end] = split.long()
        links[start + k] = (end + 8 * torch.arange(n, device=device))[:, None] + torch.arange(8, device=device)[None, :]
        boxes = torch.cat([boxes, _child_boxes(lvl[k]).view(-1, 6)], 0)
        non_leaf = torch.cat([non_leaf, torch.zeros(8 * n, dtype=torch.long, device=device)], 0)
        links = torch.cat([links, torch.zeros(8 * n, 8, dtype=torch.long, device=device)], 0)
        start, end = end, end + 8 * n
    # combine_empty (utils/octree.py:183-199), reproduced as written (its size test uses root_size / 2**depth)
    leaf_size = root[3:] / (2 ** depth)
    has_cell = (boxes[:, 3:] < leaf_size + 1e-4).all(-1)
    inner = non_leaf.bool().clone()
    for _ in range(depth):
        has_cell[inner] = has_cell[links[inner]].sum(-1).bool()
    non_leaf[inner] = has_cell.long()[inner]

    centers = boxes[:, :3] + boxes[:, 3:] * 0.5
    sdf_val, g = batched(centers, True)
    sdf_grad = g / torch.clamp(torch.norm(g, dim=-1, keepdim=True), min=1e-4)
    min_step = float((torch.ones(3) * cell_size / 2 ** depth).min()) + 1e-4
    return ops.PackedOctree(root, boxes, non_leaf, links, grid, sdf_val, sdf_grad, min_step, device)


class OctreeTracing(nn.Module):
    def __init__(self, object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=1,
                 sphere_tracing_iters=10, n_steps=100, n_rootfind_steps=8, max_iter=-1):
        super().__init__()
        self.object_bounding_sphere = object_bounding_sphere
        self.max_iter = max_iter
        self.sdf_octree = None
        self.last_counters = None
        self._net = None

    def bind(self, implicit_network):
        """Remember the network whose fused value + normal kernel builds the tree when ``generate`` is called the
        reference's way, with a bare ``lambda x: model.implicit_network(x)[:, 0]`` (train_pbr.py:403-407)."""
        self.__dict__["_net"] = implicit_network        # not a sub-module: state_dict keys stay the reference's
        return self

    def generate(self, sdf_fn, tex_sampler=None, implicit_network=None):
        """Build the octree.  ``implicit_network`` (ours) gives value+normal in one fused kernel; with a bare sdf_fn the
        normals come from autograd like the reference (utils/octree.py:595-617)."""
        box_min, box_max = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
        if tex_sampler is not None:   # model/octree_tracing.py:33-37
            v = tex_sampler.tex_sampler.vert.view(3, -1).permute(1, 0) * 0.5
            m = tex_sampler.tex_sampler.mask.view(-1) > 0.9
            box_max = [c.item() + 1e-3 for c in v[m].max(0)[0]]
            box_min = [c.item() - 1e-3 for c in v[m].min(0)[0]]
        net = implicit_network if implicit_network is not None else (
            self._net if self._net is not None else getattr(sdf_fn, "__self__", None))
        if net is not None and hasattr(net, "sdf_and_normal"):
            device = next(net.parameters()).device

            def both(x, want_grad):
                if want_grad:
                    return net.sdf_and_normal(x)
                return net.sdf(x), None
        else:
            device = torch.device("cuda")

            def both(x, want_grad):
                if not want_grad:
                    with torch.no_grad():
                        return sdf_fn(x), None
                with torch.enable_grad():
                    xx = x.detach().clone().requires_grad_(True)
                    y = sdf_fn(xx)
                    g = torch.autograd.grad(y, xx, torch.ones_like(y))[0]
                return y.detach(), g.detach()
        # The tree is built with the exact-fp32 FFMA evaluation of the SDF network: a node splits when
        # |sdf(centre)| < 0.5 |size| (utils/octree.py:381-385), so values that differ in the last bits move a few
        # borderline nodes and, through them, individual hit points by up to one march step (1e-3).  The FFMA kernel
        # reproduces the reference's tree (node count within 64, golden masks identical); the build is a one-off
        # outside the hot path, so it keeps that kernel whatever engine evaluates the network per step.
        old_engine = ops.ENGINE.get("sdf")
        ops.ENGINE["sdf"] = "ffma"
        try:
            self.sdf_octree = build_octree(both, [box_min, box_max], device)
        finally:
            ops.ENGINE["sdf"] = old_engine
        self.sdf_octree.max_iter = self.max_iter
        return self.sdf_octree

    def forward(self, sdf=None, cam_loc=None, object_mask=None, ray_directions=None):
        """cam_loc [B,3], ray_directions [B,N,3] -> points [B*N,3], mask [B*N] bool, dist [B*N]."""
        if self.sdf_octree is None:
            raise RobirError("OctreeTracing.forward called before generate()")
        B, N, _ = ray_directions.shape
        out = ops.octree_cast(self.sdf_octree, cam_loc.reshape(B, 3), ray_directions.reshape(-1, 3),
                              max_iter=self.max_iter, o_div=N, return_stats=True)
        self.last_counters = out[3]
        return out[0], out[1], out[2]
