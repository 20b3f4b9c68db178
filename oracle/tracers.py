"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  (see oracle/robir_oracle.py header for the import rules)

CPU restatement of the two surface tracers on RobIR's hot path:

  * ``OctreeOracle``  -- utils/octree.py:75-265 (Octree grid + 4 subdivision levels), :375-409 (OctreeSDF build),
    :421-438 (cast + first-order plane refinement), :459-471 (fast_volume_render micro-march), :493-585
    (multi_step_cast), model/octree_tracing.py:31-60 (generate / forward).
  * ``ray_tracing``   -- model/ray_tracing.py:26-326 (IDR two-sided sphere tracing, sampler, secant; eval mode and the
    training-only minimal-SDF branch with its uniform draw passed in).

Parity status: pinned against the unmodified reference through ``oracle/ref_shim.py``
(tests/test_oracle_vs_reference.py) and against tests/golden/*.npz.

The octree is held as plain tensors (a dict) so the same arrays can be handed to the CUDA product's packer.
All quirks listed in SURVEY.md Appendix A.3 are reproduced on purpose.
"""
import numpy as np
import torch

from robir_oracle import sphere_intersection


# ----------------------------------------------------------------------------------------------------------------------
# box helpers (utils/octree.py:10-72)
# ----------------------------------------------------------------------------------------------------------------------
def _intersect_box(boxes, o, d):
    """forward_only variant of intersect_box (:41-57): returns valid [K,1], near>=0 [K,1], far [K,1]."""
    inv = 1.0 / d
    ta = (boxes[..., :3] - o) * inv
    tb = (boxes[..., 3:] + boxes[..., :3] - o) * inv
    t1 = torch.minimum(ta, tb)
    t2 = torch.maximum(ta, tb)
    near = torch.maximum(torch.maximum(t1[..., 0:1], t1[..., 1:2]), t1[..., 2:3])
    far = torch.minimum(torch.minimum(t2[..., 0:1], t2[..., 1:2]), t2[..., 2:3])
    return torch.logical_and(near <= far, far >= 0), torch.maximum(near, torch.zeros_like(near)), far


def _inside_open(box, x):
    """inside_box(exactly=True) (:19-29): strictly inside."""
    r = (x - box[..., :3]) / box[..., 3:]
    return (r < 1).all(-1) & (r > 0).all(-1)


def _child_boxes(boxes):
    """divide (:60-72): [n,6] -> [n,8,6], child c = 4*ix + 2*iy + iz."""
    c = torch.arange(8)
    ofs = torch.stack([(c // 4) % 2, (c // 2) % 2, c % 2], -1)
    mn = boxes[:, None, :3] + ofs * boxes[:, None, 3:] / 2
    sz = (boxes[:, None, 3:] / 2).expand(mn.shape)
    return torch.cat([mn, sz], -1)


class OctreeOracle:
    """OctreeSDF + Octree + OctreeTracing (see module docstring)."""

    def __init__(self, sdf_fn, grad_fn, bounds=((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)), thr=0.5, max_iter=-1,
                 cell_size=0.05, depth=4, verbose=False):
        """sdf_fn(x[k,3]) -> [k]; grad_fn(x[k,3]) -> [k,3] (d sdf / d x).  Build: octree.py:124-199, 377-409."""
        lo = torch.tensor(bounds[0], dtype=torch.float32)
        size = torch.tensor([bounds[1][i] - bounds[0][i] for i in range(3)], dtype=torch.float32)
        cells = (size / cell_size).ceil().long()
        size = cells * cell_size  # root box rounded up to a multiple of the cell (float32 product, as the reference)
        self.root = torch.cat([lo, size])
        gx, gy, gz = [torch.arange(int(c)) for c in cells]
        anchor = torch.stack(torch.meshgrid([gx, gy, gz], indexing="ij"), -1).view(-1, 3)
        bmin = (anchor / cells) * self.root[3:] + self.root[:3]
        bmax = ((anchor + 1.0) / cells) * self.root[3:] + self.root[:3]
        boxes = torch.cat([bmin, bmax - bmin], -1)
        non_leaf = torch.zeros(boxes.shape[0], dtype=torch.long)
        links = -torch.ones(boxes.shape[0], 8, dtype=torch.long)
        self.grid = torch.arange(anchor.shape[0]).view(*[int(c) for c in cells])

        start, end = 0, boxes.shape[0]
        for _ in range(depth):
            lvl = boxes[start:end]
            with torch.no_grad():
                sz = lvl[:, 3:]
                split = sdf_fn(lvl[:, :3] + sz * 0.5).abs() < sz.norm(dim=-1) * thr
            if not split.any():
                break
            k = split.nonzero()[:, 0]
            n = k.shape[0]
            non_leaf[start:end] = split.long()
            links[start + k] = (end + 8 * torch.arange(n))[:, None] + torch.arange(8)[None, :]
            boxes = torch.cat([boxes, _child_boxes(lvl[k]).view(-1, 6)], 0)
            non_leaf = torch.cat([non_leaf, torch.zeros(8 * n, dtype=torch.long)], 0)
            links = torch.cat([links, torch.zeros(8 * n, 8, dtype=torch.long)], 0)  # add_nodes zero-fills links
            start, end = end, end + 8 * n

        # combine_empty (:183-199): internal nodes without any finest-level descendant become leaves again
        leaf_size = self.root[3:] / (2 ** depth)
        has_cell = (boxes[:, 3:] < leaf_size + 1e-4).all(-1)
        inner = non_leaf.bool().clone()
        for _ in range(depth):
            has_cell[inner] = has_cell[links[inner]].sum(-1).bool()
        non_leaf[inner] = has_cell.long()[inner]

        self.boxes, self.non_leaf, self.links = boxes, non_leaf, links
        centers = boxes[:, :3] + boxes[:, 3:] * 0.5
        chunk = 8192
        g = torch.cat([grad_fn(centers[j:j + chunk].float()).detach() for j in range(0, centers.shape[0], chunk)], 0)
        self.sdf_grad = g / torch.clamp(torch.norm(g, dim=-1, keepdim=True), min=1e-4)
        with torch.no_grad():
            self.sdf_val = torch.cat([sdf_fn(centers[j:j + chunk].float()).detach()
                                      for j in range(0, centers.shape[0], chunk)], 0)
        self.centers = centers
        self.min_step = float((torch.ones(3) * cell_size / 2 ** depth).min()) + 1e-4
        self.hit_ptr = torch.relu(self.sdf_val) <= 1e-4
        self.max_iter = max_iter
        if verbose:
            print("[octree oracle] nodes", boxes.shape[0])

    def arrays(self):
        return dict(root=self.root, boxes=self.boxes, non_leaf=self.non_leaf, links=self.links, grid=self.grid,
                    sdf_val=self.sdf_val, sdf_grad=self.sdf_grad, centers=self.centers, hit_ptr=self.hit_ptr,
                    min_step=self.min_step, max_iter=self.max_iter)

    # -- point location (Octree.query :217-265) -------------------------------------------------------------------------
    def query(self, x):
        """[k,3] -> node index [k]; -1 outside (strictly) the root box."""
        ptr_all = -torch.ones(x.shape[0], dtype=torch.long)
        inside = _inside_open(self.root, x)
        xi = x[inside]
        if xi.numel() == 0:
            return ptr_all
        res = torch.tensor(self.grid.shape)
        ijk = (((xi - self.root[:3]) / self.root[3:]) * res).floor().long()
        ptr = self.grid[ijk[:, 0], ijk[:, 1], ijk[:, 2]]
        live = self.non_leaf[ptr].nonzero()[:, 0]
        while live.numel() > 0:
            b = self.boxes[ptr[live]]
            c = torch.clip((((xi[live] - b[:, :3]) / b[:, 3:]) * 2).long(), 0, 1)
            oc = 4 * c[:, 0] + 2 * c[:, 1] + c[:, 2]
            ptr[live] = torch.gather(self.links[ptr[live]], -1, oc[:, None])[:, 0]
            live = self.non_leaf[ptr].nonzero()[:, 0]
        ptr_all[inside] = ptr
        return ptr_all

    # -- micro march over cached values (fast_volume_render :459-471) ---------------------------------------------------
    def _micro_march(self, o, d, n_samp, step):
        t = torch.linspace(0, 1, n_samp + 1) * n_samp * step + step
        pts = o[:, None, :] + d[:, None, :] * t[1:][:, None]
        ptr = self.query(pts.reshape(-1, 3))
        sdf = self.sdf_val[ptr].view(-1, n_samp)  # ptr == -1 silently reads the last node (A.3)
        hit = torch.cat([sdf <= step, torch.ones(sdf.shape[0], 1, dtype=torch.bool)], -1)
        first = torch.argmax(hit.int(), dim=-1)  # first_nonzero (:588-592)
        return t[first]

    # -- multi_step_cast (:493-585) ---------------------------------------------------------------------------------------
    def march(self, rays_o, rays_d, eps=1e-3, trace=None):
        o = rays_o.reshape(-1, 3)
        d = rays_d.reshape(-1, 3)
        if self.max_iter > 0:
            o = o + d * 0.005
        valid, near, _ = _intersect_box(self.root, o, d)
        k = valid[:, 0]
        t = near + eps
        t[~k] = -1
        ptr = -torch.ones(o.shape[0], dtype=torch.long)
        pos = torch.zeros_like(o)
        pos[k] = o[k] + t[k] * d[k]
        if k.any():
            ptr[k] = self.query(pos[k])
        k = ptr >= 0
        it = 0
        K = k.numel()
        while k.any():
            if self.max_iter > 0 and it > self.max_iter:
                break
            pk, dk = pos[k], d[k]
            _, _, far = _intersect_box(self.boxes[ptr[k]], pk, dk)
            step = 0.001
            if self.max_iter > 0:
                step = 0.01 if K > 100000 else 0.005
            live = int(k.sum())
            ms = int(np.clip(int(np.clip(K * 10, 1, 2000000) // live), 1, 100))
            if trace is not None:
                trace.append((it, live, ms))
            small = (far < ms * step)[:, 0]
            far[small] = self._micro_march(pk[small], dk[small], ms, step)[:, None]
            t[k] += far + eps
            pos[k] = o[k] + t[k] * dk
            inside = torch.ones_like(k)
            inside[k] = _inside_open(self.root, pos[k])
            k = k & inside
            ptr[~inside] = -1
            if k.any():
                ptr[k] = self.query(pos[k])
                k = ptr >= 0
            if k.any():
                kk = k.clone()
                k[kk] = ~self.hit_ptr[ptr[kk]]
            it += 1
        return t, ptr, pos

    # -- cast (:421-438) + OctreeTracing.forward (octree_tracing.py:43-60) --------------------------------------------------
    def cast(self, rays_o, rays_d):
        """-> hit_t [K], is_hit [K] bool."""
        t, ptr, x = self.march(rays_o, rays_d)
        t = t[:, 0]
        ok = ptr >= 0
        if ok.any():
            nrm = self.sdf_grad[ptr[ok]]
            foot = self.centers[ptr[ok]] - nrm * self.sdf_val[ptr[ok]].view(-1, 1)
            dist = ((foot - x[ok]) * nrm).sum(-1)
            speed = (rays_d.reshape(-1, 3)[ok] * nrm).sum(-1)
            speed[speed == 0] = 1e-4
            t[ok] += torch.clamp(dist / speed, -self.min_step * 10, self.min_step * 10)
        return t, ok

    def trace(self, cam_loc, ray_dirs):
        """OctreeTracing.forward: cam_loc [B,3], ray_dirs [B,N,3] -> points [B*N,3], mask [B*N], dist [B*N]."""
        o = cam_loc[:, None, :].expand(ray_dirs.shape).reshape(-1, 3)
        d = ray_dirs.reshape(-1, 3)
        t, ok = self.cast(o, d)
        return (t[:, None] * d + o).float(), ok, t.float()


# ----------------------------------------------------------------------------------------------------------------------
# IDR sphere tracer (model/ray_tracing.py)
# ----------------------------------------------------------------------------------------------------------------------
def ray_tracing(sdf, cam_loc, object_mask, ray_dirs, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=3,
                sphere_tracing_iters=10, n_steps=100, n_secant_steps=32, object_bounding_sphere=1.0, training=False,
                uniform_steps=None, stats=None):
    """RayTracing.forward (:26-100).  sdf: [k,3] -> [k].  cam_loc [B,3], object_mask [B*N] bool, ray_dirs [B,N,3].
    ``uniform_steps`` [n_steps] replaces the CPU ``uniform_`` draw of minimal_sdf_points (:305), training only.
    Returns points [B*N,3], network_object_mask [B*N], dists [B*N]."""
    B, N, _ = ray_dirs.shape
    flat_d = ray_dirs.reshape(-1, 3)
    flat_o = cam_loc.unsqueeze(1).expand(B, N, 3).reshape(-1, 3)

    def count(k):
        if stats is not None:
            stats["sdf_queries"] = stats.get("sdf_queries", 0) + int(k)

    def q(p):
        count(p.shape[0])
        return sdf(p)

    isect, hit_sphere = sphere_intersection(cam_loc, ray_dirs, r=object_bounding_sphere)
    isect = isect.reshape(-1, 2)
    hit_sphere = hit_sphere.reshape(-1)

    # ---- two-sided sphere tracing (:102-206)
    un_s = hit_sphere.clone()
    un_e = hit_sphere.clone()
    acc_s = torch.zeros(B * N)
    acc_e = torch.zeros(B * N)
    acc_s[un_s] = isect[un_s, 0]
    acc_e[un_e] = isect[un_e, 1]
    p_s = torch.zeros(B * N, 3)
    p_e = torch.zeros(B * N, 3)
    p_s[un_s] = (flat_o + isect[:, 0:1] * flat_d)[un_s]
    p_e[un_e] = (flat_o + isect[:, 1:2] * flat_d)[un_e]
    min_dis, max_dis = acc_s.clone(), acc_e.clone()
    nxt_s = torch.zeros(B * N)
    nxt_e = torch.zeros(B * N)
    nxt_s[un_s] = q(p_s[un_s])
    nxt_e[un_e] = q(p_e[un_e])
    it = 0
    while True:
        cur_s = torch.zeros(B * N)
        cur_s[un_s] = nxt_s[un_s]
        cur_s[cur_s <= sdf_threshold] = 0
        cur_e = torch.zeros(B * N)
        cur_e[un_e] = nxt_e[un_e]
        cur_e[cur_e <= sdf_threshold] = 0
        un_s = un_s & (cur_s > sdf_threshold)
        un_e = un_e & (cur_e > sdf_threshold)
        if (un_s.sum() == 0 and un_e.sum() == 0) or it == sphere_tracing_iters:
            break
        it += 1
        acc_s = acc_s + cur_s
        acc_e = acc_e - cur_e
        p_s = flat_o + acc_s[:, None] * flat_d
        p_e = flat_o + acc_e[:, None] * flat_d
        nxt_s = torch.zeros(B * N)
        nxt_s[un_s] = q(p_s[un_s])
        nxt_e = torch.zeros(B * N)
        nxt_e[un_e] = q(p_e[un_e])
        bad_s = nxt_s < 0
        bad_e = nxt_e < 0
        back = 0
        while (bad_s.sum() > 0 or bad_e.sum() > 0) and back < line_step_iters:
            f = (1 - line_search_step) / (2 ** back)
            acc_s[bad_s] -= f * cur_s[bad_s]
            p_s[bad_s] = (flat_o + acc_s[:, None] * flat_d)[bad_s]
            acc_e[bad_e] += f * cur_e[bad_e]
            p_e[bad_e] = (flat_o + acc_e[:, None] * flat_d)[bad_e]
            nxt_s[bad_s] = q(p_s[bad_s])
            nxt_e[bad_e] = q(p_e[bad_e])
            bad_s = nxt_s < 0
            bad_e = nxt_e < 0
            back += 1
        un_s = un_s & (acc_s < acc_e)
        un_e = un_e & (acc_s < acc_e)

    net_mask = acc_s < acc_e
    points = p_s
    dists = acc_s

    # ---- sampler + secant for the non-convergent rays (:208-297)
    samp = un_s
    if samp.sum() > 0:
        idx = torch.nonzero(samp).flatten()
        lin = torch.linspace(0, 1, steps=n_steps).view(1, -1)
        z = acc_s[idx, None] + lin * (acc_e[idx] - acc_s[idx])[:, None]          # [m, n_steps]
        pts = flat_o[idx, None, :] + z[:, :, None] * flat_d[idx, None, :]
        vals = torch.cat([q(c) for c in torch.split(pts.reshape(-1, 3), 100000, dim=0)]).reshape(-1, n_steps)
        rank = torch.sign(vals) * torch.arange(n_steps, 0, -1).float().reshape(1, n_steps)
        first = torch.argmin(rank, -1)
        ar = torch.arange(idx.shape[0])
        s_pts = pts[ar, first]
        s_z = z[ar, first]
        true_surf = object_mask.reshape(-1)[idx]
        net_surf = vals[ar, first] < 0
        p_out = ~(true_surf & net_surf)
        if p_out.sum() > 0:
            amin = torch.argmin(vals[p_out], -1)
            s_pts[p_out] = pts[p_out][torch.arange(int(p_out.sum())), amin]
            s_z[p_out] = z[p_out][torch.arange(int(p_out.sum())), amin]
        samp_net = samp.clone()
        samp_net[idx[~net_surf]] = False
        sec = (net_surf & true_surf) if training else net_surf
        if sec.sum() > 0:
            z_hi = z[ar, first][sec]
            f_hi = vals[ar, first][sec]
            z_lo = z[sec][torch.arange(int(sec.sum())), first[sec] - 1]
            f_lo = vals[sec][torch.arange(int(sec.sum())), first[sec] - 1]
            so, sdir = flat_o[idx[sec]], flat_d[idx[sec]]
            zp = (-f_lo * (z_hi - z_lo) / (f_hi - f_lo + 1e-8) + z_lo).clamp(0.0, 2e1)
            for _ in range(n_secant_steps):
                fm = q(so + zp[:, None] * sdir)
                lo = fm > 0
                z_lo = torch.where(lo, zp, z_lo)
                f_lo = torch.where(lo, fm, f_lo)
                hi = fm < 0
                z_hi = torch.where(hi, zp, z_hi)
                f_hi = torch.where(hi, fm, f_hi)
                zp = (-f_lo * (z_hi - z_lo) / (f_hi - f_lo + 1e-8) + z_lo).clamp(0.0, 2e1)
            s_pts[sec] = so + zp[:, None] * sdir
            s_z[sec] = zp
        points = points.clone()
        dists = dists.clone()
        points[idx] = s_pts
        dists[idx] = s_z
        net_mask = net_mask.clone()
        net_mask[idx] = samp_net[idx]

    if not training:
        return points, net_mask, dists

    # ---- training only: minimal-SDF points for non-hit rays (:73-100, 299-326)
    om = object_mask.reshape(-1)
    in_mask = ~net_mask & om & ~samp
    out_mask = ~om & ~samp
    left = (in_mask | out_mask) & ~hit_sphere
    if left.sum() > 0:
        dists[left] = -(flat_d[left] * flat_o[left]).sum(-1)
        points[left] = flat_o[left] + dists[left].unsqueeze(1) * flat_d[left]
    m = (in_mask | out_mask) & hit_sphere
    if m.sum() > 0:
        sel = net_mask & out_mask
        min_dis[sel] = dists[sel]
        steps = uniform_steps.unsqueeze(0) * (max_dis[m] - min_dis[m]).unsqueeze(-1) + min_dis[m].unsqueeze(-1)
        cand = flat_o[m].unsqueeze(1) + steps.unsqueeze(-1) * flat_d[m].unsqueeze(1)
        vals = torch.cat([q(c) for c in torch.split(cand.reshape(-1, 3), 100000, dim=0)]).reshape(-1, n_steps)
        amin = vals.argmin(-1)
        ar = torch.arange(int(m.sum()))
        points[m] = cand[ar, amin]
        dists[m] = steps[ar, amin]
    return points, net_mask, dists
