"""Python-side operators over the C ABI: tensor allocation, weight packing caches and torch.autograd.Function
wrappers whose forward/backward are the hand-written CUDA kernels.  PyTorch only owns memory, streams and the
autograd graph between operators."""
import ctypes
import math

import torch

from . import _lib
from ._lib import (MlpParams, OctCastParams, OctreeView, SdfParams, SdfTcParams, SgParams, SphereTraceParams, TlParams, check, f32, lib, ptr,
                   sm_count, stream)

TINY = 1e-6


def _empty(*shape, dtype=torch.float32, like=None):
    return torch.empty(*shape, dtype=dtype, device=like.device)


def _zeros(*shape, dtype=torch.float32, like=None):
    return torch.zeros(*shape, dtype=dtype, device=like.device)


# ----------------------------------------------------------------------------------------------------------------------
# weight packing (cached on parameter versions; weights that are being trained are re-packed when they change)
# ----------------------------------------------------------------------------------------------------------------------
_pack_epoch = [0]


def invalidate_packed_weights():
    """Call after an in-place parameter update that does not bump the tensors' version counters (e.g.
    torch.optim.Adam(fused=True): its multi-tensor kernel leaves ``_version`` untouched), so that the packed copies of
    trainable weights are rebuilt on the next use.  Optimizers that go through ordinary in-place ops (the reference's
    torch.optim.Adam default) need nothing."""
    _pack_epoch[0] += 1


class _PackCache:
    """Packed / transposed / tensor-core images of a set of parameters, keyed on their storage and version counters.
    While a CUDA graph is being captured, parameters that are being trained are always re-packed: the replayed graph
    must contain the packing kernels whatever the cache state was at capture time (the optimizer runs inside the same
    graph)."""

    def __init__(self):
        self.key = None
        self.val = None
        self.trained = False

    def get(self, tensors, build):
        # "being trained" = has received a gradient at some point (an optimizer may have updated it in place since the
        # last use); sticky, because zero_grad(set_to_none=True) clears the gradients between forward and backward
        self.trained = self.trained or any(t.requires_grad and t.grad is not None for t in tensors)
        trainable = self.trained
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors) + \
            ((_pack_epoch[0],) if trainable else ())
        # Safety net for callers that capture a graph without ever announcing parameter updates: parameters that
        # have received a gradient are re-packed on every use during capture.  GraphedPBRStep announces every step
        # (invalidate_packed_weights at the top of the captured region), which makes the key change exactly once.
        capturing = trainable and _pack_epoch[0] == 0 and tensors[0].is_cuda and \
            torch.cuda.is_current_stream_capturing()
        if key != self.key or capturing:
            self.val = build()
            self.key = key
        return self.val


def pack_transpose(W, k_begin, k_count, Kpad, Npad, scale=1.0):
    W = f32(W)
    N, K = W.shape
    out = _empty(Kpad, Npad, like=W)
    check(lib().robir_pack_transpose(ptr(W), N, K, k_begin, k_count, ptr(out), Kpad, Npad, scale, stream()))
    return out


def pack_window(W, k_begin, k_count, Npad, Kpad):
    W = f32(W)
    N, K = W.shape
    out = _empty(Npad, Kpad, like=W)
    check(lib().robir_pack_window(ptr(W), N, K, k_begin, k_count, ptr(out), Npad, Kpad, stream()))
    return out


def pack_wn_transpose(v, g, n_begin, n_count, Kpad, Npad):
    v, g = f32(v), f32(g)
    N, K = v.shape
    out = _empty(Kpad, Npad, like=v)
    check(lib().robir_pack_wn_transpose(ptr(v), ptr(g), N, K, n_begin, n_count, ptr(out), Kpad, Npad, stream()))
    return out


class VisWeights:
    """Packed VisNetwork weights (implicit_differentiable_renderer.py:225-258: 126 -> 256 x4 -> 2, ReLU)."""

    def __init__(self, vis_layer):
        self.layers = [vis_layer[i] for i in (0, 2, 4, 6, 8)]
        self.cache = _PackCache()

    def get(self):
        L = self.layers
        tensors = [t for l in L for t in (l.weight, l.bias)]

        def build():
            W0 = L[0].weight
            if tuple(W0.shape) != (256, 126) or any(tuple(L[i].weight.shape) != (256, 256) for i in (1, 2, 3)) \
                    or tuple(L[4].weight.shape) != (2, 256):
                raise _lib.RobirError("VisNetwork must be 126->256->256->256->256->2 (points/dirs multires 10)")
            d = {}
            d["Wt0p"] = pack_transpose(W0, 0, 63, 64, 256)
            d["Wt0d"] = pack_transpose(W0, 63, 63, 64, 256)
            d["W0d"] = pack_window(W0, 63, 63, 256, 256)
            d["b0"] = f32(L[0].bias)
            for i in (1, 2, 3):
                d["Wt%d" % i] = pack_transpose(L[i].weight, 0, 256, 256, 256)
                d["W%d" % i] = f32(L[i].weight)
                d["b%d" % i] = f32(L[i].bias)
            W4, b4 = f32(L[4].weight), f32(L[4].bias)
            d["wd"] = (W4[1] - W4[0]).contiguous()
            d["bd"] = (b4[1] - b4[0]).reshape(1).contiguous()
            # tensor-core engines: pre-swizzled, power-of-two-scaled fp16 weight images in streaming order
            # (csrc/vis_tc.cu); terms = 3 -> hi/lo (fp32 parity, engine "tc"), terms = 1 -> hi only (engine "tc1")
            W0c = f32(W0)
            for terms in (3, 1):
                sb = lib().robir_tc_image_bytes(1, 0, terms)
                fwd = torch.empty(lib().robir_tc_image_bytes(3, 0, terms), dtype=torch.uint8, device=W0.device)
                bwd = torch.zeros(lib().robir_tc_image_bytes(3, 1, terms), dtype=torch.uint8, device=W0.device)
                for j, i in enumerate((1, 2, 3)):
                    tc_pack_layer(d["W%d" % i], 256, 256, 256, False, 2, fwd[j * sb:], terms)
                for j, i in enumerate((3, 2, 1)):
                    tc_pack_layer(d["W%d" % i], 256, 256, 256, True, 2, bwd[j * sb:], terms)
                tc_pack_layer(W0c.reshape(-1)[63:], 126, 63, 256, True, 1, bwd[3 * sb:], terms)
                d["tc_fwd%d" % terms], d["tc_bwd%d" % terms] = fwd, bwd
            d["bias3"] = torch.stack([d["b1"], d["b2"], d["b3"]]).contiguous()
            return d
        return self.cache.get(tensors, build)


def vis_engine_terms(engine=None):
    """MMA terms per logical product of the visibility engine: 3 = fp32-parity split ("tc"), 1 = fast mode ("tc1")."""
    engine = ENGINE["vis"] if engine is None else engine
    return {"tc": 3, "tc1": 1}.get(engine, 0)


def tc_pack_layer(W, ldw, N, K, transpose, n_halves, out, terms=3):
    check(lib().robir_tc_pack_layer(ptr(W), ldw, N, K, int(transpose), n_halves, terms, ptr(out), stream()))


def tc_selftest(A, W, terms=3):
    """D[128,256] = A[128,256] @ W[256,256]^T through the tcgen05 machinery (terms = 3: scaled fp16 hi/lo split, fp32
    parity; terms = 1: single-pass fp16) -- unit test hook."""
    A, W = f32(A), f32(W)
    img = torch.empty(lib().robir_tc_image_bytes(1, 0, terms), dtype=torch.uint8, device=A.device)
    tc_pack_layer(W, 256, 256, 256, False, 2, img, terms)
    D = torch.zeros(128, 256, device=A.device)
    check(lib().robir_tc_selftest(ptr(A), ptr(img), ptr(D), terms, stream()))
    return D


_point_tab = {"key": None, "tab": None}


class point_table_scope:
    """Inside the scope the layer-0 point table W0[:, :63] PE(p) + b0 is computed once per (points, weights) and shared
    by the diffuse and the BRDF-lobe visibility queries of one render_with_all_sg call."""

    def __enter__(self):
        _point_tab["key"], _point_tab["tab"], self.on = None, None, True
        _point_tab["on"] = True
        return self

    def __exit__(self, *exc):
        _point_tab["key"], _point_tab["tab"], _point_tab["on"] = None, None, False
        return False


def point_table(W, points):
    if not _point_tab.get("on"):
        return pe_linear(points, W["Wt0p"], W["b0"])
    key = (points.data_ptr(), tuple(points.shape), id(W))
    if _point_tab["key"] != key:
        _point_tab["key"], _point_tab["tab"] = key, pe_linear(points, W["Wt0p"], W["b0"])
    return _point_tab["tab"]


def pe_linear(x, Wt, bias):
    x = f32(x)
    n = x.shape[0]
    out = _empty(n, 256, like=x)
    check(lib().robir_pe_linear(ptr(x), n, ptr(Wt), ptr(bias), ptr(out), stream()))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# sample directions
# ----------------------------------------------------------------------------------------------------------------------
class _SampleDirs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axis_f, axis_w, sharp, lam_w, sg_range, u_theta, u_phi, renorm):
        K, S = u_theta.shape
        axis_f, axis_w, sharp, lam_w, sg_range = map(f32, (axis_f, axis_w, sharp, lam_w, sg_range))
        u_theta, u_phi = f32(u_theta), f32(u_phi)
        dirs = _empty(K * S, 3, like=axis_f)
        w = _empty(K * S, like=axis_f)
        check(lib().robir_sample_dirs_fwd(K, S, ptr(axis_f), ptr(axis_w), ptr(sharp), ptr(lam_w), ptr(sg_range),
                                          ptr(u_theta), ptr(u_phi), int(renorm), ptr(dirs), ptr(w), stream()))
        ctx.save_for_backward(axis_f, axis_w, sharp, lam_w, sg_range, u_theta, u_phi)
        ctx.renorm = int(renorm)
        return dirs, w

    @staticmethod
    def backward(ctx, g_dirs, g_w):
        axis_f, axis_w, sharp, lam_w, sg_range, u_theta, u_phi = ctx.saved_tensors
        K, S = u_theta.shape
        g_dirs = f32(g_dirs) if g_dirs is not None else _zeros(K * S, 3, like=axis_f)
        g_w = f32(g_w) if g_w is not None else _zeros(K * S, like=axis_f)
        g_af = _empty(K, 3, like=axis_f)
        g_aw = _zeros(K, 3, like=axis_f)
        g_sharp = _empty(K, like=axis_f)
        g_lam = _empty(K, like=axis_f)
        g_rng = _zeros(1, like=axis_f)
        check(lib().robir_sample_dirs_bwd(K, S, ptr(axis_f), ptr(axis_w), ptr(sharp), ptr(lam_w), ptr(sg_range),
                                          ptr(u_theta), ptr(u_phi), ctx.renorm, ptr(g_dirs), ptr(g_w), ptr(g_af),
                                          None if ctx.renorm else ptr(g_aw), ptr(g_sharp), ptr(g_lam), ptr(g_rng),
                                          stream()))
        return g_af, g_aw, g_sharp, g_lam, g_rng.reshape(sg_range.shape), None, None, None


def sample_dirs(axis_f, axis_w, sharp, lam_w, sg_range, u_theta, u_phi, renorm):
    return _SampleDirs.apply(axis_f, axis_w, sharp, lam_w, sg_range, u_theta, u_phi, renorm)


# ----------------------------------------------------------------------------------------------------------------------
# fused visibility queries
# ----------------------------------------------------------------------------------------------------------------------
class Stats:
    """Device-side counters of executed work (for the algorithmic-FLOP roofline): executed
    visibility queries of the per-lobe diffuse list ([0]) and of the BRDF-lobe lists ([1]), accumulated over calls."""
    n_pairs = None          # int64 device tensor [2]

    @classmethod
    def pairs_tensor(cls, like, kind=0):
        if cls.n_pairs is None or cls.n_pairs.device != like.device:
            cls.n_pairs = torch.zeros(2, dtype=torch.int64, device=like.device)
        return cls.n_pairs[kind:kind + 1]

    @classmethod
    def reset(cls):
        if cls.n_pairs is not None:
            cls.n_pairs.zero_()

    @classmethod
    def total(cls):
        return int(cls.n_pairs.sum().item()) if cls.n_pairs is not None else 0

    @classmethod
    def diffuse(cls):
        return int(cls.n_pairs[0].item()) if cls.n_pairs is not None else 0


class active_rows:
    """Context: a device int32 scalar n_act; inside it the per-point networks (fused MLP chains, SDF) evaluate only
    rows [0, n_act) of their fixed-capacity batch (hit rays compacted to the front) and write zeros for the rest."""
    current = None

    def __init__(self, n_act):
        self.n_act = n_act

    def __enter__(self):
        self.old, active_rows.current = active_rows.current, self.n_act
        return self

    def __exit__(self, *exc):
        active_rows.current = self.old
        return False


# vis: visibility MLP; mlp: the 512-wide encoder / lobe chains of the material and indirect-illumination networks.
# vis: "tc": tcgen05, scaled fp16 hi/lo 3-term split (fp32 parity, default) | "tc1": tcgen05 single-pass fp16 (fast mode,
# ~1e-4: NOT the parity mode) | "ffma": exact-fp32 CUDA cores.  mlp: "tc" (bf16 hi/lo layer engine) | "ffma"
# wn: the CESR stage's weight-normed 512-wide chains (shadow_net / normal_net): "tc" | "torch" (cuBLAS cross-check)
# sdf: the NeuS SDF network (value / normal / features): "tc" (tcgen05, csrc/sdf_tc.cu) | "ffma" (csrc/sdf.cu)
ENGINE = {"vis": "tc", "mlp": "tc", "wn": "tc", "sdf": "tc"}
PROFILE = None             # when a list: (name, start_event, end_event, max_tiles) per hot-kernel launch (bench.py)


class _Timed:
    def __init__(self, name, max_tiles):
        self.name, self.max_tiles = name, max_tiles

    def __enter__(self):
        if PROFILE is not None:
            self.s, self.e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e.record()
            PROFILE.append((self.name, self.s, self.e, self.max_tiles))
        return False


def tile_rows():
    return 128 if ENGINE["vis"] in ("tc", "tc1") else 64


def _vis_mlp_fwd(W, tabA, tabB, rowA, rowB, n_tiles, max_tiles, need_mask):
    rows = max_tiles * tile_rows()
    vis = _empty(rows, like=tabA)
    mask = _empty(rows, 4, 8, dtype=torch.int32, like=tabA) if need_mask else None
    terms = vis_engine_terms()
    if terms:
        with _Timed("vis_mlp_fwd", max_tiles):
            check(lib().robir_vis_tc_fwd(ptr(tabA), ptr(tabB), ptr(rowA), ptr(rowB), ptr(n_tiles), max_tiles,
                                         ptr(W["tc_fwd%d" % terms]), ptr(W["bias3"]), ptr(W["wd"]), ptr(W["bd"]),
                                         ptr(vis), ptr(mask), terms, sm_count(), stream()))
        return vis, mask
    with _Timed("vis_mlp_fwd", max_tiles):
        check(lib().robir_vis_mlp_fwd(ptr(tabA), ptr(tabB), ptr(rowA), ptr(rowB), ptr(n_tiles), max_tiles,
                                      ptr(W["Wt1"]), ptr(W["Wt2"]), ptr(W["Wt3"]), ptr(W["b1"]), ptr(W["b2"]),
                                      ptr(W["b3"]), ptr(W["wd"]), ptr(W["bd"]), ptr(vis), ptr(mask), sm_count(),
                                      stream()))
    return vis, mask


# SMs the big visibility backward leaves free: the backward of the material / indirect networks does not depend on it
# and runs on other streams at the same time, but its kernels cannot share an SM with a visibility CTA (shared memory);
# with every SM taken they would queue behind it and then stretch its tail.  The kernel draws tiles from a counter, so
# a smaller persistent grid only costs its share of the throughput.
VIS_BWD_RESERVED_SMS = 12


def _vis_mlp_bwd(W, rowB, n_tiles, max_tiles, vis, g_vis, mask, dirs, engine):
    g_dirs = _zeros(dirs.shape[0], 3, like=dirs)
    terms = vis_engine_terms(engine)
    if terms:
        ctas = sm_count() - (VIS_BWD_RESERVED_SMS if max_tiles > 4 * sm_count() else 0)
        with _Timed("vis_mlp_bwd", max_tiles):
            check(lib().robir_vis_tc_bwd(ptr(rowB), ptr(n_tiles), max_tiles, ptr(W["tc_bwd%d" % terms]), ptr(W["wd"]),
                                         ptr(vis), ptr(g_vis), ptr(mask), ptr(dirs), ptr(g_dirs), terms, max(ctas, 1),
                                         stream()))
        return g_dirs
    with _Timed("vis_mlp_bwd", max_tiles):
        check(lib().robir_vis_mlp_bwd(ptr(rowB), ptr(n_tiles), max_tiles, ptr(W["W1"]), ptr(W["W2"]), ptr(W["W3"]),
                                      ptr(W["W0d"]), ptr(W["wd"]), ptr(vis), ptr(g_vis), ptr(mask), ptr(dirs),
                                      ptr(g_dirs), sm_count(), stream()))
    return g_dirs


class _DiffuseVis(torch.autograd.Function):
    """light_vis[n, M] = per-lobe weighted mean of the visibility MLP over the S sample directions of each lobe
    (model/sg_render.py:148-195).  Differentiable w.r.t. dirs and w only."""

    @staticmethod
    def forward(ctx, points, normals, dirs, w, M, S, weights, need_grad, tabB=None):
        W = weights.get()
        points, normals, dirs, w = map(f32, (points, normals, dirs, w))
        n = points.shape[0]
        T = tile_rows()
        cap = ((M * S + T - 1) // T) * T
        dev = points
        bits = _empty(n, M, dtype=torch.int32, like=dev)
        lobe_off = _empty(n, M + 1, dtype=torch.int32, like=dev)
        start = _empty(n, dtype=torch.int32, like=dev)
        rowA = _empty(n * cap, dtype=torch.int32, like=dev)
        rowB = _empty(n * cap, dtype=torch.int32, like=dev)
        n_tiles = _zeros(1, dtype=torch.int32, like=dev)
        # tensor-core engine: points packed back to back (rows of one tile may belong to two points)
        check(lib().robir_diffuse_rows(n, M, S, T, 0 if vis_engine_terms() else 1, ptr(normals), ptr(dirs), ptr(bits),
                                       ptr(lobe_off), ptr(start),
                                       ptr(rowA), ptr(rowB), ptr(n_tiles), ptr(Stats.pairs_tensor(dev)), stream()))
        tabA = point_table(W, points)
        if tabB is None:
            tabB = pe_linear(dirs, W["Wt0d"], None)
        max_tiles = n * cap // T
        vis, mask = _vis_mlp_fwd(W, tabA, tabB, rowA, rowB, n_tiles, max_tiles, need_grad)
        lv = _empty(n, M, like=dev)
        check(lib().robir_diffuse_reduce_fwd(n, M, S, ptr(bits), ptr(lobe_off), ptr(start), ptr(vis), ptr(w), ptr(lv),
                                             stream()))
        if need_grad:
            ctx.save_for_backward(dirs, w, bits, lobe_off, start, rowB, n_tiles, vis, mask, lv)
            ctx.meta = (n, M, S, max_tiles, weights, ENGINE["vis"])
        return lv

    @staticmethod
    def backward(ctx, g_lv):
        dirs, w, bits, lobe_off, start, rowB, n_tiles, vis, mask, lv = ctx.saved_tensors
        n, M, S, max_tiles, weights, engine = ctx.meta
        W = weights.get()
        g_lv = f32(g_lv)
        g_vis = _zeros(vis.shape[0], like=vis)
        g_w = _zeros(M * S, like=vis)
        check(lib().robir_diffuse_reduce_bwd(n, M, S, ptr(bits), ptr(lobe_off), ptr(start), ptr(vis), ptr(w), ptr(lv),
                                             ptr(g_lv), ptr(g_vis), ptr(g_w), stream()))
        g_dirs = _vis_mlp_bwd(W, rowB, n_tiles, max_tiles, vis, g_vis, mask, dirs, engine)
        return None, None, g_dirs, g_w, None, None, None, None, None


class _SpecVis(torch.autograd.Function):
    """brdf_vis[n] = weighted mean of the visibility MLP over S per-point sample directions
    (model/sg_render.py:242-301, single view).  inv = None: the batch holds the direct (rows [0, n)) and the indirect
    (rows [n, 2n), class 0 = "inv") call of render_with_all_sg over the same n points -> out [2n]."""

    @staticmethod
    def forward(ctx, points, normals, dirs, w, S, inv, testing, weights, need_grad):
        W = weights.get()
        points, normals, dirs, w = map(f32, (points, normals, dirs, w))
        n = points.shape[0]
        copies = 2 if inv is None else 1
        nq = n * copies
        T = tile_rows()
        rows = ((nq * S + T - 1) // T) * T
        dev = points
        rowA = _empty(rows, dtype=torch.int32, like=dev)
        rowB = _empty(rows, dtype=torch.int32, like=dev)
        n_tiles = _zeros(1, dtype=torch.int32, like=dev)
        check(lib().robir_spec_rows(nq, S, rows, T, n if copies > 1 else 0, ptr(normals), ptr(dirs), ptr(rowA),
                                    ptr(rowB), ptr(n_tiles), ptr(Stats.pairs_tensor(dev, 1)), stream()))
        tabA = point_table(W, points)
        tabB = pe_linear(dirs, W["Wt0d"], None)
        vis, mask = _vis_mlp_fwd(W, tabA, tabB, rowA, rowB, n_tiles, rows // T, need_grad)
        out = _empty(nq, like=dev)
        for c in range(copies):
            o = c * n * S
            check(lib().robir_spec_reduce_fwd(n, S, int(inv) if inv is not None else c, int(testing),
                                              c_ptr(rowB, o * 4), c_ptr(vis, o * 4), c_ptr(w, o * 4),
                                              c_ptr(out, c * n * 4), stream()))
        if need_grad:
            ctx.save_for_backward(dirs, w, rowB, n_tiles, vis, mask, out)
            ctx.meta = (n, S, inv, rows, weights, T, ENGINE["vis"])
        return out

    @staticmethod
    def backward(ctx, g_out):
        dirs, w, rowB, n_tiles, vis, mask, out = ctx.saved_tensors
        n, S, inv, rows, weights, T, engine = ctx.meta
        copies = 2 if inv is None else 1
        W = weights.get()
        g_out = f32(g_out)
        g_vis = _zeros(rows, like=vis)
        g_w = _empty(n * copies * S, like=vis)
        for c in range(copies):
            o = c * n * S
            check(lib().robir_spec_reduce_bwd(n, S, int(inv) if inv is not None else c, c_ptr(rowB, o * 4),
                                              c_ptr(vis, o * 4), c_ptr(w, o * 4), c_ptr(out, c * n * 4),
                                              c_ptr(g_out, c * n * 4), c_ptr(g_vis, o * 4), c_ptr(g_w, o * 4), stream()))
        g_dirs = _vis_mlp_bwd(W, rowB, n_tiles, rows // T, vis, g_vis, mask, dirs, engine)
        return None, None, g_dirs, g_w, None, None, None, None, None


class _SpecPrep(torch.autograd.Function):
    """Sampling inputs shared by the direct and the indirect BRDF-lobe visibility call (csrc/vis.cu spec_prep_*):
    (normal, view, roughness [n,1], valid) -> ref [2n,3], wl [2n,3], sharp [2n], sg_range [1].  Only the roughness is
    differentiable (render_with_all_sg passes detached normals / view directions, sg_render.py:324-327)."""

    @staticmethod
    def forward(ctx, normal, view, rough, valid):
        normal, view, r = f32(normal), f32(view), f32(rough).reshape(-1)
        n = normal.shape[0]
        ref, wl = _empty(2 * n, 3, like=normal), _empty(2 * n, 3, like=normal)
        sharp, sg_range, wlam = _empty(2 * n, like=normal), _empty(1, like=normal), _empty(n, like=normal)
        argmin = _empty(1, dtype=torch.int32, like=normal)
        v8 = valid.to(torch.uint8).contiguous() if valid is not None else None
        check(lib().robir_spec_prep_fwd(n, ptr(normal), ptr(view), ptr(r), ptr(v8), ptr(ref), ptr(wl), ptr(sharp),
                                        ptr(sg_range), ptr(wlam), ptr(argmin), stream()))
        ctx.save_for_backward(r, wlam, argmin)
        ctx.rshape = tuple(rough.shape)
        ctx.mark_non_differentiable(ref, wl)
        return ref, wl, sharp, sg_range

    @staticmethod
    def backward(ctx, _g_ref, _g_wl, g_sharp, g_range):
        r, wlam, argmin = ctx.saved_tensors
        n = r.shape[0]
        g_sharp = f32(g_sharp) if g_sharp is not None else _zeros(2 * n, like=r)
        g_range = f32(g_range) if g_range is not None else _zeros(1, like=r)
        g_rough = _empty(n, like=r)
        check(lib().robir_spec_prep_bwd(n, ptr(r), ptr(wlam), ptr(argmin), ptr(g_sharp), ptr(g_range), ptr(g_rough),
                                        stream()))
        return None, None, g_rough.reshape(ctx.rshape), None


def spec_prep(normal, view, rough, valid=None):
    return _SpecPrep.apply(normal, view, rough, valid)


def diffuse_vis(points, normals, dirs, w, M, S, weights, need_grad, tabB=None):
    """tabB: optional pre-computed direction table pe_linear(dirs, Wt0d) of the same weights."""
    return _DiffuseVis.apply(points, normals, dirs, w, M, S, weights, need_grad, tabB)


def spec_vis(points, normals, dirs, w, S, inv, testing, weights, need_grad):
    return _SpecVis.apply(points, normals, dirs, w, S, inv, testing, weights, need_grad)


# ----------------------------------------------------------------------------------------------------------------------
# SG render
# ----------------------------------------------------------------------------------------------------------------------
class _SgRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normal, view, rough, albedo, spec_refl, lgt, ind_lgt, light_vis, bv_dir, bv_ind, ind_integral,
                lin_diff):
        normal, view, rough, albedo, spec_refl, lgt, light_vis, bv_dir = map(
            f32, (normal, view, rough, albedo, spec_refl, lgt, light_vis, bv_dir))
        n, M = normal.shape[0], lgt.shape[0]
        has_ind = ind_lgt is not None
        Mi = ind_lgt.shape[1] if has_ind else 0
        if has_ind:
            ind_lgt, bv_ind, ind_integral = f32(ind_lgt), f32(bv_ind), f32(ind_integral)
        outs = [_empty(n, 3, like=normal) for _ in range(7)]
        pre = _empty(n, 9, like=normal)
        p = SgParams()
        p.n, p.M, p.Mi, p.lin_diff = n, M, Mi, int(lin_diff)
        for k, t in dict(normal=normal, view=view, rough=rough, albedo=albedo, spec_refl=spec_refl, lgt=lgt,
                         ind_lgt=ind_lgt if has_ind else None, light_vis=light_vis, bv_dir=bv_dir,
                         bv_ind=bv_ind if has_ind else bv_dir, ind_integral=ind_integral if has_ind else None,
                         sg_rgb=outs[0], sg_spec=outs[1], sg_diff=outs[2], vis_shadow=outs[3], ind_rgb=outs[4],
                         ind_spec=outs[5], ind_diff=outs[6], pre=pre).items():
            setattr(p, k, ptr(t))
        check(lib().robir_sg_render_fwd(ctypes.byref(p), stream()))
        ctx.save_for_backward(normal, view, rough, albedo, spec_refl, lgt, ind_lgt if has_ind else None, light_vis,
                              bv_dir, bv_ind if has_ind else None, ind_integral if has_ind else None, pre)
        ctx.meta = (n, M, Mi, int(lin_diff), tuple(rough.shape))
        ctx.mark_non_differentiable(outs[3])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_rgb, g_spec, g_diff, _g_shadow, g_irgb, g_ispec, g_idiff):
        normal, view, rough, albedo, spec_refl, lgt, ind_lgt, light_vis, bv_dir, bv_ind, ind_integral, pre = \
            ctx.saved_tensors
        n, M, Mi, lin_diff, rough_shape = ctx.meta
        has_ind = Mi > 0
        z = lambda *s: _zeros(*s, like=normal)
        g_lgt, g_ind_lgt, g_lv = z(M, 7), (z(n, Mi, 7) if has_ind else None), z(n, M)
        g_bvd, g_bvi, g_rough, g_alb, g_sr, g_int = z(n), z(n), z(n), z(n, 3), z(1), z(n, 3)
        g_nrm = z(n, 3) if ctx.needs_input_grad[0] else None      # shading normal: only the CESR stage asks for it
        p = SgParams()
        p.n, p.M, p.Mi, p.lin_diff = n, M, Mi, lin_diff
        fz = lambda g: f32(g) if g is not None else None
        for k, t in dict(normal=normal, view=view, rough=rough, albedo=albedo, spec_refl=spec_refl, lgt=lgt,
                         ind_lgt=ind_lgt, light_vis=light_vis, bv_dir=bv_dir, bv_ind=bv_ind if has_ind else bv_dir,
                         ind_integral=ind_integral, pre=pre, g_sg_rgb=fz(g_rgb), g_sg_spec=fz(g_spec),
                         g_sg_diff=fz(g_diff), g_ind_rgb=fz(g_irgb), g_ind_spec=fz(g_ispec), g_ind_diff=fz(g_idiff),
                         g_lgt=g_lgt, g_ind_lgt=g_ind_lgt, g_light_vis=g_lv, g_bv_dir=g_bvd, g_bv_ind=g_bvi,
                         g_rough=g_rough, g_albedo=g_alb, g_spec_refl=g_sr, g_ind_integral=g_int,
                         g_normal=g_nrm).items():
            setattr(p, k, ptr(t))
        check(lib().robir_sg_render_bwd(ctypes.byref(p), stream()))
        return (g_nrm, None, g_rough.reshape(rough_shape), g_alb, g_sr.reshape(spec_refl.shape), g_lgt, g_ind_lgt, g_lv,
                g_bvd, g_bvi if has_ind else None, g_int if has_ind else None, None)


def sg_render(normal, view, rough, albedo, spec_refl, lgt, ind_lgt, light_vis, bv_dir, bv_ind, ind_integral,
              lin_diff=False):
    return _SgRender.apply(normal, view, rough, albedo, spec_refl, lgt, ind_lgt, light_vis, bv_dir, bv_ind,
                           ind_integral, lin_diff)


def compact_hits(hit, points, dirs):
    """Stable hit-first partition of a ray batch on the device (no host sync): -> pos [N] int64 (slot of ray i),
    order [N] int64 (ray in slot), n_act [1] int32, valid [N] bool (slot < n_act), pts [N,3] (hit points in slot order,
    0 for the misses), view [N,3] (= -dirs in slot order).  No autograd (points / dirs come from the no-grad tracer)."""
    hit8 = hit.to(torch.uint8).contiguous()
    points, dirs = f32(points), f32(dirs)
    N = hit8.shape[0]
    pos = _empty(N, dtype=torch.int64, like=points)
    order = _empty(N, dtype=torch.int64, like=points)
    n_act = _empty(1, dtype=torch.int32, like=points)
    valid = _empty(N, dtype=torch.uint8, like=points)
    pts, view = _empty(N, 3, like=points), _empty(N, 3, like=points)
    check(lib().robir_compact_hits(N, ptr(hit8), ptr(points), ptr(dirs), ptr(pos), ptr(order), ptr(n_act), ptr(valid),
                                   ptr(pts), ptr(view), stream()))
    return pos, order, n_act, valid.view(torch.bool), pts, view


class _LatentPair(torch.autograd.Function):
    """[sigmoid(z); sigmoid(z) + 0.01 noise] -> [2n, 32]  (SparseAE.forward with smooth_on_latent, latent_dim 32)"""

    @staticmethod
    def forward(ctx, z, noise):
        z, noise = f32(z), f32(noise)
        n = z.shape[0]
        out = _empty(2 * n, 32, like=z)
        check(lib().robir_latent_pair_fwd(n, ptr(z), ptr(noise), ptr(out), stream()))
        ctx.save_for_backward(z)
        return out

    @staticmethod
    def backward(ctx, g):
        z, = ctx.saved_tensors
        g_z = _empty(*z.shape, like=z)
        g = f32(g)                  # a contiguous copy (if one is made) must outlive the launch: bind it to a name
        check(lib().robir_latent_pair_bwd(z.shape[0], ptr(z), ptr(g), ptr(g_z), stream()))
        return g_z, None


class _BrdfHead(torch.autograd.Function):
    """decoder outputs [2n, 5] -> albedo [n,3], roughness [n,1], metallic [n,1] and their random_xi twins"""

    @staticmethod
    def forward(ctx, y2):
        y2 = f32(y2)
        n = y2.shape[0] // 2
        outs = [_empty(n, 3, like=y2), _empty(n, 1, like=y2), _empty(n, 1, like=y2),
                _empty(n, 3, like=y2), _empty(n, 1, like=y2), _empty(n, 1, like=y2)]
        check(lib().robir_brdf_head_fwd(n, ptr(y2), *[ptr(o) for o in outs], stream()))
        ctx.save_for_backward(y2)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        y2, = ctx.saved_tensors
        g_y2 = _empty(*y2.shape, like=y2)
        # upstream gradients may be strided views (slices of a concatenated gradient): f32() then makes contiguous
        # copies, and ALL of them have to stay alive until the launch -- a temporary that dies inside the argument list
        # hands its block to the next copy, and two gradients end up at the same address
        gs = [f32(g) if g is not None else None for g in gs]
        check(lib().robir_brdf_head_bwd(y2.shape[0] // 2, ptr(y2), *[ptr(g) for g in gs], ptr(g_y2), stream()))
        return g_y2


def latent_pair(z, noise):
    return _LatentPair.apply(z, noise)


def brdf_head(y2):
    return _BrdfHead.apply(y2)


class _DecodeLobes(torch.autograd.Function):
    """IndirctIllumNetwork._decode_lobes (implicit_differentiable_renderer.py:207-219) in one launch each way."""

    @staticmethod
    def forward(ctx, raw):
        raw = f32(raw)
        n, L = raw.shape[0], raw.shape[1]
        sgs = _empty(n, L, 7, like=raw)
        check(lib().robir_decode_lobes_fwd(n * L, ptr(raw), ptr(sgs), stream()))
        ctx.save_for_backward(raw)
        return sgs

    @staticmethod
    def backward(ctx, g):
        raw, = ctx.saved_tensors
        g_raw = _empty(*raw.shape, like=raw)
        g = f32(g)
        check(lib().robir_decode_lobes_bwd(raw.shape[0] * raw.shape[1], ptr(raw), ptr(g), ptr(g_raw), stream()))
        return g_raw


def decode_lobes(raw):
    """raw [n, lobes, 6] -> SGs [n, lobes, 7]"""
    return _DecodeLobes.apply(raw)


# ----------------------------------------------------------------------------------------------------------------------
# SDF network
# ----------------------------------------------------------------------------------------------------------------------
class SdfWeights:
    """Folded + packed SDFNetwork weights (model/neus_model.py:312-438)."""

    def __init__(self, sdf_network):
        self.net = sdf_network
        self.cache = _PackCache()

    def get(self):
        net = self.net
        lins = [getattr(net, "lin%d" % l) for l in range(9)]
        tensors = [t for l in lins for t in (l.weight_v, l.weight_g, l.bias)]

        def build():
            d = {}
            for l in range(8):
                v, g = lins[l].weight_v, lins[l].weight_g
                N, K = v.shape
                exp = (256, 63) if l == 0 else ((193, 256) if l == 3 else (256, 256))
                if (N, K) != exp:
                    raise _lib.RobirError("SDFNetwork layer %d has shape %s, expected %s" % (l, (N, K), exp))
                d["Wt%d" % l] = pack_wn_transpose(v, g, 0, N, 64 if l == 0 else 256, 256)
                b = torch.zeros(256, device=v.device)
                b[:N] = f32(lins[l].bias)
                d["b%d" % l] = b
            v8, g8 = lins[8].weight_v, lins[8].weight_g
            if tuple(v8.shape) != (257, 256):
                raise _lib.RobirError("SDFNetwork last layer must be 256 -> 257")
            d["Wt8_feat"] = pack_wn_transpose(v8, g8, 1, 256, 256, 256)
            v8c, g8c = f32(v8), f32(g8)
            w8 = _empty(256, like=v8c)
            check(lib().robir_pack_wn_row(ptr(v8c), ptr(g8c), 256, 0, ptr(w8), stream()))
            d["w8_sdf"] = w8
            d["b8"] = f32(lins[8].bias)
            # tensor-core engine (csrc/sdf_tc.cu): scaled fp16 hi/lo images of the nine 256-wide layers in streaming
            # order; the 1/sqrt(2) of the skip concat is folded into layer 4
            sb = lib().robir_tc_image_bytes(1, 0, 3)
            img = torch.zeros(9 * sb, dtype=torch.uint8, device=w8.device)
            for l in range(8):
                Wt = d["Wt%d" % l] if l != 4 else (d["Wt4"] * (1.0 / math.sqrt(2.0))).contiguous()
                tc_pack_layer(Wt, 256, 256, 64 if l == 0 else 256, True, 2, img[l * sb:], 3)
            tc_pack_layer(d["Wt8_feat"], 256, 256, 256, True, 2, img[8 * sb:], 3)
            d["tc_img"] = img
            d["bias8"] = torch.stack([d["b%d" % l] for l in range(8)]).contiguous()
            return d
        return self.cache.get(tensors, build)


def sdf_eval(weights, pts, in_scale=2.0, sdf_scale=0.5, feat_scale=0.5, want_grad=False, want_feat=False):
    """-> sdf [n], grad [n,3] | None, feat [n,256] | None   (no autograd: the SDF network is frozen in stage 2)."""
    W = weights.get()
    pts = f32(pts)
    n = pts.shape[0]
    # measured against the FFMA kernel (tools/c3_profile.py): 3.3x / 6.0x at 740 k points with / without the normal and
    # features (158 / 241 TFLOP/s algorithmic), 1.2x at the 550 hit points of a training batch; below a few tiles both
    # are latency-bound and the FFMA kernel's 32-row tiles spread better
    if ENGINE["sdf"] == "tc" and n * (4 if want_grad else 1) >= SDF_TC_MIN_ROWS:
        return _sdf_eval_tc(W, pts, in_scale, sdf_scale, feat_scale, want_grad, want_feat)
    sdf = _empty(n, like=pts)
    grad = _empty(n, 3, like=pts) if want_grad else None
    feat = _empty(n, 256, like=pts) if want_feat else None
    p = SdfParams()
    p.pts, p.n, p.in_scale, p.sdf_scale, p.feat_scale = ptr(pts), n, in_scale, sdf_scale, feat_scale
    for l in range(8):
        p.Wt[l] = W["Wt%d" % l].data_ptr()
        p.bias[l] = W["b%d" % l].data_ptr()
    p.w8_sdf, p.b8, p.Wt8_feat = ptr(W["w8_sdf"]), ptr(W["b8"]), ptr(W["Wt8_feat"])
    p.sdf, p.grad, p.feat = ptr(sdf), ptr(grad), ptr(feat)
    p.n_active = ptr(active_rows.current)
    check(lib().robir_sdf_eval(ctypes.byref(p), sm_count(), stream()))
    return sdf, grad, feat


SDF_TC_MIN_ROWS = 128 * 8


def _sdf_eval_tc(W, pts, in_scale, sdf_scale, feat_scale, want_grad, want_feat):
    n = pts.shape[0]
    alloc = _zeros if active_rows.current is not None else _empty      # inactive rows must read as 0
    sdf = alloc(n, like=pts)
    grad = alloc(n, 3, like=pts) if want_grad else None
    feat = alloc(n, 256, like=pts) if want_feat else None
    p = SdfTcParams()
    p.pts, p.n, p.in_scale, p.sdf_scale, p.feat_scale = ptr(pts), n, in_scale, sdf_scale, feat_scale
    p.img, p.bias, p.w8_sdf, p.b8 = ptr(W["tc_img"]), ptr(W["bias8"]), ptr(W["w8_sdf"]), ptr(W["b8"])
    p.sdf, p.grad, p.feat = ptr(sdf), ptr(grad), ptr(feat)
    p.n_active = ptr(active_rows.current)
    check(lib().robir_sdf_tc(ctypes.byref(p), sm_count(), stream()))
    return sdf, grad, feat


# ----------------------------------------------------------------------------------------------------------------------
# camera rays / octree
# ----------------------------------------------------------------------------------------------------------------------
def _pose_from_quaternion(pose7):
    """[1,7] = unit-normalised quaternion (r, i, j, k) + camera position -> [1,4,4] camera-to-world matrix
    (utils/rend_util.py:52-57,100-117: the Hamilton-convention rotation matrix of the normalised quaternion)."""
    q = torch.nn.functional.normalize(pose7[:, :4].float(), dim=1)
    r, i, j, k = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [1 - 2 * (j * j + k * k), 2 * (j * i - k * r), 2 * (i * k + r * j),
            2 * (j * i + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r),
            2 * (k * i - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)]
    p = torch.eye(4, device=pose7.device, dtype=torch.float32).repeat(pose7.shape[0], 1, 1)
    p[:, :3, :3] = torch.stack(rows, -1).reshape(-1, 3, 3)
    p[:, :3, 3] = pose7[:, 4:].float()
    return p


def camera_rays(uv, pose, intrinsics):
    """uv [1,N,2], pose [1,4,4] (or [1,7]: quaternion + position), intrinsics [1,3,3] (focal lengths, principal point,
    skew) -> ray_dirs [1,N,3], cam_loc [1,3]  (utils/rend_util.py:51-97)."""
    if pose.dim() == 2 and pose.shape[1] == 7:
        pose = _pose_from_quaternion(pose)
    if uv.shape[0] != 1 or pose.shape[1:] != (4, 4):
        raise _lib.RobirError("camera_rays supports one pose (4x4 matrix or 7-vector) per call (the reference's "
                              "training/eval usage)")
    uv, pose, K = f32(uv), f32(pose), f32(intrinsics)
    N = uv.shape[1]
    dirs = _empty(1, N, 3, like=uv)
    check(lib().robir_camera_rays(N, ptr(uv), ptr(pose), ptr(K), ptr(dirs), stream()))
    return dirs, pose[:, :3, 3].contiguous()


class PackedOctree:
    """Device copy of the cached-SDF octree in the 32-byte node layout of csrc/octree_walk.h."""

    def __init__(self, root, boxes, non_leaf, links, grid, sdf_val, sdf_grad, min_step, device):
        n = boxes.shape[0]
        nodes = torch.empty(n, 8, dtype=torch.float32)
        nodes[:, :6] = boxes.float().cpu()
        child = torch.where(non_leaf.cpu().reshape(-1) != 0, links.cpu()[:, 0], torch.full((n,), -1, dtype=torch.long))
        nodes[:, 6] = child.to(torch.int32).view(torch.float32)
        nodes[:, 7] = sdf_val.float().cpu()
        self.nodes = nodes.to(device).contiguous()
        self.grid = grid.to(torch.int32).to(device).contiguous()
        self.sdf_grad = sdf_grad.float().to(device).contiguous()
        self.root = [float(v) for v in root.float().cpu()]
        self.n_nodes = n
        self.last_sdf = float(sdf_val.float().cpu()[-1])
        self.min_step = float(min_step)
        # torch.clamp(x, -min_step*10, min_step*10): python-double product, then cast to fp32 (utils/octree.py:433)
        self.refine_limit = float(torch.tensor(self.min_step * 10, dtype=torch.float32))
        self.node_bytes = self.nodes.numel() * 4 + self.grid.numel() * 4 + self.sdf_grad.numel() * 4

    def host_arrays(self):
        """Unpack to the plain-tensor layout of the reference / oracle octree (CPU), e.g. for bench.py's CPU leg."""
        nodes = self.nodes.cpu()
        boxes = nodes[:, :6].contiguous()
        child = nodes[:, 6].contiguous().view(torch.int32).long()
        non_leaf = (child >= 0).long()
        links = torch.where(child[:, None] >= 0, child[:, None] + torch.arange(8)[None, :], torch.zeros(1, dtype=torch.long))
        sdf_val = nodes[:, 7].contiguous()
        return dict(root=torch.tensor(self.root), boxes=boxes, non_leaf=non_leaf, links=links,
                    grid=self.grid.cpu().long(), sdf_val=sdf_val, sdf_grad=self.sdf_grad.cpu(),
                    centers=boxes[:, :3] + boxes[:, 3:] * 0.5, hit_ptr=torch.relu(sdf_val) <= 1e-4,
                    min_step=self.min_step)

    def view(self):
        v = OctreeView()
        v.nodes, v.grid = ptr(self.nodes), ptr(self.grid)
        v.gx, v.gy, v.gz = self.grid.shape
        v.n_nodes = self.n_nodes
        v.rminx, v.rminy, v.rminz, v.rsizex, v.rsizey, v.rsizez = self.root
        return v


def octree_cast(tree, rays_o, rays_d, max_iter=-1, o_div=1, return_stats=False):
    """OctreeSDF.cast + OctreeTracing.forward (utils/octree.py:421-438, model/octree_tracing.py:43-60).
    rays_o [K/o_div, 3], rays_d [K,3] -> hit_x [K,3], is_hit [K] bool, hit_t [K]."""
    rays_o, rays_d = f32(rays_o), f32(rays_d)
    K = rays_d.shape[0]
    dev = rays_d
    out_t, out_x = _empty(K, like=dev), _empty(K, 3, like=dev)
    out_hit = _empty(K, dtype=torch.uint8, like=dev)
    if K == 0:
        if return_stats:
            return out_x, out_hit.bool(), out_t, _zeros(lib().robir_octree_counters_len(), dtype=torch.int32, like=dev)
        return out_x, out_hit.bool(), out_t
    st_t, st_p = _empty(K, like=dev), _empty(K, dtype=torch.int32, like=dev)
    counters = _zeros(lib().robir_octree_counters_len(), dtype=torch.int32, like=dev)
    p = OctCastParams()
    p.view = tree.view()
    p.sdf_grad, p.rays_o, p.rays_d = ptr(tree.sdf_grad), ptr(rays_o), ptr(rays_d)
    p.K, p.o_div, p.max_iter, p.eps = K, o_div, max_iter, 1e-3
    p.refine_limit, p.last_node_sdf = tree.refine_limit, tree.last_sdf
    p.state_t, p.state_ptr, p.out_t, p.out_x, p.out_hit, p.counters = map(ptr, (st_t, st_p, out_t, out_x, out_hit,
                                                                                counters))
    check(lib().robir_octree_cast(ctypes.byref(p), sm_count(), stream()))
    if return_stats:
        return out_x, out_hit.bool(), out_t, counters
    return out_x, out_hit.bool(), out_t


def sphere_trace(weights, cam_loc, ray_dirs, object_mask, radius=1.0, sdf_threshold=5.0e-5, line_search_step=0.5,
                 line_step_iters=1, sphere_tracing_iters=10, n_steps=100, n_secant_steps=8, training=False,
                 uniform_steps=None, in_scale=2.0, out_scale=0.5, return_stats=False):
    """RayTracing.forward (model/ray_tracing.py:26-100).  cam_loc [B,3], ray_dirs [B,N,3], object_mask [B*N] bool|None
    -> points [B*N,3], network_object_mask [B*N] bool, dists [B*N]  (+ int32 counters [8] with return_stats)."""
    W = weights.get()
    B, Np, _ = ray_dirs.shape
    cam_loc, dirs = f32(cam_loc).reshape(B, 3), f32(ray_dirs).reshape(-1, 3)
    K = B * Np
    dev = dirs
    points, dists = _empty(K, 3, like=dev), _empty(K, like=dev)
    mask = _empty(K, dtype=torch.uint8, like=dev)
    counters = _zeros(8, dtype=torch.int32, like=dev)
    if K == 0:
        return (points, mask.bool(), dists, counters) if return_stats else (points, mask.bool(), dists)
    om = object_mask.reshape(-1).to(torch.uint8).contiguous() if object_mask is not None else None
    fws = _empty(8, K, like=dev)                         # acc_s, acc_e, min_dis, max_dis, sec_state[K][4]
    iws = _empty(3, K, dtype=torch.int32, like=dev)      # samp_list, sec_list, min_list
    flags = _empty(K, dtype=torch.uint8, like=dev)
    vals = _empty(K * n_steps, like=dev)
    p = SphereTraceParams()
    for l in range(8):
        p.net.Wt[l] = W["Wt%d" % l].data_ptr()
        p.net.bias[l] = W["b%d" % l].data_ptr()
    p.net.w8_sdf, p.net.b8 = ptr(W["w8_sdf"]), ptr(W["b8"])
    p.N, p.o_div, p.cam_loc, p.ray_dirs, p.object_mask = K, Np, ptr(cam_loc), ptr(dirs), ptr(om)
    p.in_scale, p.out_scale, p.radius, p.sdf_threshold = in_scale, out_scale, radius, sdf_threshold
    p.line_search_step, p.line_step_iters, p.sphere_tracing_iters = line_search_step, line_step_iters, sphere_tracing_iters
    p.n_steps, p.n_secant_steps, p.training = n_steps, n_secant_steps, int(bool(training))
    if training:
        if uniform_steps is None or uniform_steps.numel() != n_steps:
            raise _lib.RobirError("sphere_trace(training=True) needs uniform_steps[n_steps]")
        uniform_steps = f32(uniform_steps).to(dev.device)
        p.uniform_steps = ptr(uniform_steps)
    p.points, p.net_mask, p.dists = ptr(points), ptr(mask), ptr(dists)
    p.acc_s, p.acc_e, p.min_dis, p.max_dis = (c_ptr(fws, i * K * 4) for i in range(4))
    p.sec_state = c_ptr(fws, 4 * K * 4)
    p.samp_list, p.sec_list, p.min_list = (c_ptr(iws, i * K * 4) for i in range(3))
    p.flags, p.vals, p.counters = ptr(flags), ptr(vals), ptr(counters)
    check(lib().robir_sphere_trace(ctypes.byref(p), sm_count(), stream()))
    if return_stats:
        return points, mask.bool(), dists, counters
    return points, mask.bool(), dists


def c_ptr(t, byte_offset):
    return ctypes.c_void_p(t.data_ptr() + byte_offset)


# ----------------------------------------------------------------------------------------------------------------------
# fused small MLP chains (material / indirect networks)
# ----------------------------------------------------------------------------------------------------------------------
ACT = {"none": 0, "relu": 1, "leaky": 2, "softplus100": 3}     # enum Act of csrc/common.cuh
IN_MODE = {"raw": 0, "pe10": 1, "pe10_extra": 2, "ipe10": 3, "pe10x2": 4}


def _up(x, m):
    return (x + m - 1) // m * m


class MlpChain:
    """A chain of nn.Linear layers (an nn.Sequential of the reference modules) evaluated by one fused kernel.
    Packed copies (transposed for the forward, padded for the backward) are cached per parameter version."""

    def __init__(self, linears, acts, in_mode):
        self.linears, self.acts, self.in_mode = list(linears), [ACT[a] for a in acts], IN_MODE[in_mode]
        assert len(self.linears) == len(self.acts) <= 8
        self.cache = _PackCache()

    @classmethod
    def from_sequential(cls, seq, hidden_act, in_mode):
        lins = [m for m in seq if isinstance(m, torch.nn.Linear)]
        return cls(lins, [hidden_act] * (len(lins) - 1) + ["none"], in_mode)

    @classmethod
    def from_weightnorm(cls, lins, hidden_act, in_mode):
        """Chain over weight-normed layers (weight_g / weight_v / bias, e.g. the NeuS colour network); forward only."""
        ch = cls([], [], in_mode)
        ch.linears, ch.acts, ch.weightnorm = list(lins), [ACT[hidden_act]] * (len(lins) - 1) + [ACT["none"]], True
        return ch

    weightnorm = False
    wgrad_streams = None      # side streams of this chain's weight-gradient GEMMs (created on first backward)

    def tc_ok(self, n):
        """The tensor-core layer engine takes chains with a 512-wide hidden layer (the encoders and the lobe network);
        the narrow decoders stay on the FFMA chain kernel."""
        if ENGINE["mlp"] != "tc" or self.weightnorm or self.in_mode == 0 or n < 128:
            return False
        dims = [(l.weight.shape[0], l.weight.shape[1]) for l in self.linears]
        return max(d[0] for d in dims) >= 256 and all(d[0] <= 512 and d[1] <= 512 for d in dims)

    def packed_tc(self):
        """Per layer: forward image of W and backward image of W^T (csrc/tc_mlp.cu), cached like packed()."""
        if "_tc_cache" not in self.__dict__:
            self.__dict__["_tc_cache"] = _PackCache()
        blk = 32768

        def build():
            out = []
            for lin in self.linears:
                W = f32(lin.weight)
                N, K = W.shape
                cbn, nkbk = (N + 127) // 128, (K + 63) // 64          # forward: rows N, contraction K
                cbk, nkbn = (K + 127) // 128, (N + 63) // 64          # backward: rows K, contraction N
                fw = torch.empty(cbn * nkbk * blk, dtype=torch.uint8, device=W.device)
                bw = torch.empty(cbk * nkbn * blk, dtype=torch.uint8, device=W.device)
                check(lib().robir_tl_pack_weight(ptr(W), K, N, K, 0, cbn, nkbk, ptr(fw), stream()))
                check(lib().robir_tl_pack_weight(ptr(W), K, K, N, 1, cbk, nkbn, ptr(bw), stream()))
                out.append(dict(fw=fw, bw=bw, nkb_fw=nkbk, nkb_bw=nkbn))
            return out
        return self.__dict__["_tc_cache"].get(self.params(), build)

    def eval_tc_ok(self, n, x):
        """Evaluation-only forward of a raw-input chain on the layer engine: large batches outside autograd (the colour
        network of borrow_color, 131 072 rows per call) -- weight-normed chains included, the weight norm is folded into
        the images."""
        if ENGINE["mlp"] != "tc" or self.in_mode != 0 or n < TL_BIG_MIN_ROWS or torch.is_grad_enabled():
            return False
        dims = [tuple((l.weight_v if self.weightnorm else l.weight).shape) for l in self.linears]
        return x.shape[1] == dims[0][1] and max(d[0] for d in dims) >= 256 and all(d[0] <= 512 and d[1] <= 512 for d in dims)

    def packed_eval_tc(self):
        """Forward images of the (folded) weights + zero-padded biases, cached on the parameters."""
        if "_tc_eval_cache" not in self.__dict__:
            self.__dict__["_tc_eval_cache"] = _PackCache()

        def build():
            out = []
            for lin in self.linears:
                if self.weightnorm:
                    v, g = f32(lin.weight_v), f32(lin.weight_g)
                    W = (g * v / v.norm(dim=1, keepdim=True)).contiguous()
                else:
                    W = f32(lin.weight)
                N, K = W.shape
                cbn, nkb = (N + 127) // 128, (K + 63) // 64
                fw = torch.empty(cbn * nkb * 32768, dtype=torch.uint8, device=W.device)
                check(lib().robir_tl_pack_weight(ptr(W), K, N, K, 0, cbn, nkb, ptr(fw), stream()))
                bias = torch.zeros(cbn * 128, device=W.device)
                bias[:N] = f32(lin.bias)
                out.append(dict(fw=fw, bias=bias, N=N, K=K, nkb=nkb))
            return out
        return self.__dict__["_tc_eval_cache"].get(self.params(), build)

    def eval_tc(self, x):
        """x [n, K0] -> [n, N_last]: one layer-engine launch per Linear, hidden activations only as hi/lo images."""
        x = f32(x)
        n = x.shape[0]
        layers = self.packed_eval_tc()
        tiles = (n + 127) // 128
        img = _tl_rows_image(x, layers[0]["K"])
        out = None
        for l, d in enumerate(layers):
            last = l == len(layers) - 1
            out = _empty(n, d["N"], like=x) if last else None
            nkb_out = 0 if last else (d["N"] + 63) // 64
            nxt = _tl_image(tiles, nkb_out, x) if nkb_out else None
            q = _tl_params(img, d["fw"], d["bias"], n, d["N"], d["nkb"], 0, self.acts[l], None, out, nxt, nkb_out, None, 0)
            _tl_layer(q, n)
            img = nxt
        return out

    def wgrad_tickets(self, layer, tiles, like):
        """Persistent zero-initialised ticket counters of robir_mlp_wgrad's in-kernel split reduction (it re-zeroes them)."""
        t = self.__dict__.setdefault("_tickets", {})
        if layer not in t or t[layer].numel() < tiles or t[layer].device != like.device:
            t[layer] = torch.zeros(tiles, dtype=torch.int32, device=like.device)
        return t[layer]

    def params(self):
        if self.weightnorm:
            return [t for l in self.linears for t in (l.weight_v, l.weight_g, l.bias)]
        return [t for l in self.linears for t in (l.weight, l.bias)]

    def packed(self):
        def build():
            if not self.weightnorm and len(self.linears) <= 8:
                # one launch for all layers (csrc/mlp.cu pack_chain_kernel): the re-pack of a trained chain is on the
                # critical path of the step once the octree walk is pipelined
                q = _lib.PackChainParams()
                q.n_layers = len(self.linears)
                out, keep = [], []
                for l, lin in enumerate(self.linears):
                    W, b = f32(lin.weight), f32(lin.bias)
                    N, K = W.shape
                    Kp, Np = _up(K, 16), _up(N, 256)
                    Wt, Wb, bp = _empty(Kp, Np, like=W), _empty(_up(N, 16), _up(K, 256), like=W), _empty(Np, like=W)
                    L = q.L[l]
                    L.W, L.b, L.Wt, L.Wb, L.bias = W.data_ptr(), b.data_ptr(), Wt.data_ptr(), Wb.data_ptr(), bp.data_ptr()
                    L.N, L.K, L.Kp, L.Np, L.Nb, L.Kb = N, K, Kp, Np, _up(N, 16), _up(K, 256)
                    keep.append((W, b))
                    out.append(dict(Wt=Wt, Wb=Wb, bias=bp, K=K, N=N, Kpad=Kp, Npad=Np))
                check(lib().robir_pack_chain(ctypes.byref(q), stream()))
                return out
            out = []
            for lin in self.linears:
                if self.weightnorm:
                    v, g = f32(lin.weight_v), f32(lin.weight_g)
                    W = g * v / v.norm(dim=1, keepdim=True)
                else:
                    W = f32(lin.weight)
                b = f32(lin.bias)
                N, K = W.shape
                Kp, Np = _up(K, 16), _up(N, 256)
                Wt = pack_transpose(W, 0, K, Kp, Np)
                Wb = _empty(_up(N, 16), _up(K, 256), like=W)
                check(lib().robir_pack_pad(ptr(W), N, K, ptr(Wb), _up(N, 16), _up(K, 256), stream()))
                bp = torch.zeros(Np, device=W.device)
                bp[:N] = b
                out.append(dict(Wt=Wt, Wb=Wb, bias=bp, K=K, N=N, Kpad=Kp, Npad=Np))
            return out
        return self.cache.get(self.params(), build)


def _mlp_params(chain, packed, n, x, extra, noise, noise_scale, n_active=None, segments=1):
    p = MlpParams()
    p.n, p.n_layers, p.in_mode = n, len(packed), chain.in_mode
    p.in_dim, p.in_pad = packed[0]["K"], packed[0]["Kpad"]
    p.x, p.extra, p.noise, p.noise_scale = ptr(x), ptr(extra), ptr(noise), float(noise_scale)
    p.n_active, p.seg = ptr(n_active), n // segments
    for l, d in enumerate(packed):
        L = p.L[l]
        L.Wt, L.Wb, L.bias = d["Wt"].data_ptr(), d["Wb"].data_ptr(), d["bias"].data_ptr()
        L.K, L.N, L.Kpad, L.Npad, L.act = d["K"], d["N"], d["Kpad"], d["Npad"], chain.acts[l]
    return p


def _tl_params(a_img, w_img, bias, n, N, nkb, mode, act, ref, out, out_img, nkb_out, n_active, seg):
    q = TlParams()
    q.a_img, q.w_img, q.bias, q.n, q.N, q.nkb, q.mode, q.act = ptr(a_img), ptr(w_img), ptr(bias), n, N, nkb, mode, act
    q.ref, q.ld_ref = ptr(ref), (ref.shape[1] if ref is not None else 0)
    q.out, q.ld_out = ptr(out), (out.shape[1] if out is not None else 0)
    q.out_img, q.nkb_out, q.n_active, q.seg = ptr(out_img), nkb_out, ptr(n_active), seg
    return q


def _tl_image(tiles, nkb, like):
    return torch.empty(tiles * nkb * 32768, dtype=torch.uint8, device=like.device)


def _tl_forward(chain, packed, tc, p, n, x0, saves, out, n_active, segments):
    """Forward of a chain on the tensor-core layer engine (csrc/tc_mlp.cu): embedding -> image -> one launch per layer.
    saves[l] (post-activation rows of the hidden layers) are written when the caller allocated them."""
    tiles, seg = (n + 127) // 128, n // segments
    seg_ok = seg % 128 == 0
    na = n_active if seg_ok else None
    check(lib().robir_mlp_encode(ctypes.byref(p), sm_count(), stream()))
    img = _tl_image(tiles, tc[0]["nkb_fw"], x0)
    check(lib().robir_tl_pack_rows(ptr(x0), x0.shape[1], n, packed[0]["K"], None, 0, 0, tc[0]["nkb_fw"], ptr(img), stream()))
    L = len(packed)
    for l, d in enumerate(packed):
        last = l == L - 1
        dst = out if last else (saves[l] if saves else None)
        nkb_out = tc[l + 1]["nkb_fw"] if not last else 0
        nxt = _tl_image(tiles, nkb_out, x0) if not last else None
        q = _tl_params(img, tc[l]["fw"], d["bias"], n, d["N"], tc[l]["nkb_fw"], 0, chain.acts[l], None, dst, nxt, nkb_out,
                       na, seg)
        _tl_layer(q, n)
        img = nxt


def _tl_backward(chain, packed, tc, n, g_out, saves, Gs, g_x, n_active, segments, want_param_grad):
    """Input-gradient chain on the layer engine: G_{l-1} = (G_l W_l) act'(A_{l-1}); Gs[l] (fp32, for the weight
    gradients) are filled when requested; g_x receives the gradient of the embedded input."""
    tiles, seg = (n + 127) // 128, n // segments
    na = n_active if seg % 128 == 0 else None
    L = len(packed)
    n_out = packed[-1]["N"]
    if want_param_grad:
        Gs[L - 1][:, :n_out].copy_(g_out)
    img = _tl_image(tiles, tc[L - 1]["nkb_bw"], g_out)
    check(lib().robir_tl_pack_rows(ptr(g_out), g_out.shape[1], n, n_out, None, 0, 0, tc[L - 1]["nkb_bw"], ptr(img), stream()))
    for l in range(L - 1, -1, -1):
        d = packed[l]
        first = l == 0
        dst = g_x if first else (Gs[l - 1] if want_param_grad else None)
        nkb_out = tc[l - 1]["nkb_bw"] if not first else 0
        nxt = _tl_image(tiles, nkb_out, g_out) if not first else None
        q = _tl_params(img, tc[l]["bw"], None, n, d["K"], tc[l]["nkb_bw"], 1, chain.acts[l - 1] if not first else 0,
                       saves[l - 1] if not first else None, dst, nxt, nkb_out, na, seg)
        _tl_layer(q, n)
        img = nxt


# ----------------------------------------------------------------------------------------------------------------------
# weight-normed softplus chains with a skip concat (CESR shadow_net / normal_net) on the tensor-core layer engine
# ----------------------------------------------------------------------------------------------------------------------
# rows from which the CESR chains' weight gradients run on the tensor cores (robir_tl_wgrad); below it the fp32 FFMA
# kernel of csrc/mlp.cu is launch-bound anyway.  Tests lower it to cover the kernel at small sizes.
WN_TC_WGRAD_MIN_ROWS = 4096
# rows from which a layer runs as persistent CTAs over 128 x 256 tiles (robir_tl_layer_big) instead of one CTA per
# 128 x 128 tile (robir_tl_layer): identical results
TL_BIG_MIN_ROWS = 4096


def _tl_layer(q, rows):
    if rows >= TL_BIG_MIN_ROWS:
        check(lib().robir_tl_layer_big(ctypes.byref(q), sm_count(), stream()))
    else:
        check(lib().robir_tl_layer(ctypes.byref(q), stream()))


def _tl_weight_images(W, rows_bw):
    """Forward image of W [N,K] and backward image of W^T restricted to its first rows_bw rows (the columns of W that
    lead back to the previous layer; the rest belongs to the skip input, which needs no gradient)."""
    N, K = W.shape
    cbn, nkbk = (N + 127) // 128, (K + 63) // 64
    fw = torch.empty(cbn * nkbk * 32768, dtype=torch.uint8, device=W.device)
    check(lib().robir_tl_pack_weight(ptr(W), K, N, K, 0, cbn, nkbk, ptr(fw), stream()))
    bw = None
    if rows_bw > 0:
        cbk, nkbn = (rows_bw + 127) // 128, (N + 63) // 64
        bw = torch.empty(cbk * nkbn * 32768, dtype=torch.uint8, device=W.device)
        check(lib().robir_tl_pack_weight(ptr(W), K, rows_bw, N, 1, cbk, nkbn, ptr(bw), stream()))
    return fw, bw


def _tl_rows_image(X, K):
    n = X.shape[0]
    nkb = (K + 63) // 64
    img = _tl_image((n + 127) // 128, nkb, X)
    check(lib().robir_tl_pack_rows(ptr(X), X.shape[1], n, K, None, 0, 0, nkb, ptr(img), stream()))
    return img


class _WnChain(torch.autograd.Function):
    """y = lin_{L-1}( ... softplus100(lin_l(a_l)) ... ), a_l = cat([h_{l-1}, x]) / sqrt(2) at the skip layers
    (SDFNetwork.forward with multires = 0, model/neus_model.py:397-415), one csrc/tc_mlp.cu launch per layer.
    The 1/sqrt(2) is folded into the skip layer's weight; the concat is a 512-column buffer that the previous layer
    writes its columns into.  backward: input-gradient chain on the same engine, dW / db by robir_tl_wgrad (tensor cores,
    R >= WN_TC_WGRAD_MIN_ROWS) or robir_mlp_wgrad."""

    @staticmethod
    def forward(ctx, x, skip, L, n_active, *wb):
        Ws, bs = wb[:L], wb[L:]
        if x.requires_grad:
            raise _lib.RobirError("wn_chain: gradients with respect to the input rows are not produced")
        x = f32(x)
        R, d_in = x.shape
        tiles = (R + 127) // 128
        SOFTPLUS = ACT["softplus100"]
        need_bwd = any(ctx.needs_input_grad)      # evaluation (plots, testing=True): keep nothing, free layer by layer
        We, imgs, a_rows, a_imgs = [], [], [], []
        a, img = x, _tl_rows_image(x, d_in)
        out = None
        for l in range(L):
            W = f32(Ws[l]) * (1.0 / math.sqrt(2.0)) if l in skip else f32(Ws[l])
            N, K = W.shape
            if K > 512 or N > 512 or a.shape[1] != K:
                raise _lib.RobirError("wn_chain: layer %d is %dx%d on %d input columns (engine limit 512)" % (l, N, K, a.shape[1]))
            fw, bw = _tl_weight_images(W, 0 if (l == 0 or not need_bwd) else (K - d_in if l in skip else K))
            if need_bwd:
                We.append(W)
                imgs.append((fw, bw))
                a_rows.append(a)
                if R >= WN_TC_WGRAD_MIN_ROWS:
                    a_imgs.append(img)          # the layer's input image doubles as the weight-gradient GEMM's operand
            bias = _zeros(((N + 127) // 128) * 128, like=x)
            bias[:N] = f32(bs[l])
            last, next_skip = l == L - 1, (l + 1) in skip
            if next_skip:
                out = _empty(R, N + d_in, like=x)                   # [h_l | x]: the next layer's input rows
                out[:, N:] = x
            else:
                out = _empty(R, N, like=x)
            nkb_out = 0 if (last or next_skip) else (N + 63) // 64
            nxt = _tl_image(tiles, nkb_out, x) if nkb_out else None
            q = _tl_params(img, fw, bias, R, N, (K + 63) // 64, 0, 0 if last else SOFTPLUS, None, out, nxt, nkb_out,
                           n_active, 0)
            # hidden layers of a fixed-capacity batch: every consumer of these rows (next layer, backward, tensor-core
            # weight gradients) skips the row tiles beyond n_active, so they are not zero-filled (189 MB per layer)
            q.no_fill = int(n_active is not None and not last and R >= max(TL_BIG_MIN_ROWS, WN_TC_WGRAD_MIN_ROWS))
            _tl_layer(q, R)
            a = out
            img = _tl_rows_image(out, N + d_in) if next_skip else nxt
        ctx.meta = (R, d_in, L, tuple(skip))
        ctx.imgs = imgs
        ctx.a_imgs = a_imgs
        ctx.n_active = n_active
        ctx.save_for_backward(*a_rows, *We)
        return out

    @staticmethod
    def backward(ctx, g_out):
        R, d_in, L, skip = ctx.meta
        a_rows, We = ctx.saved_tensors[:L], ctx.saved_tensors[L:]
        tiles = (R + 127) // 128
        SOFTPLUS = ACT["softplus100"]
        n_active = ctx.n_active
        G = g0 = f32(g_out)              # g0: device / dtype reference for allocations (G may become None below)
        img = _tl_rows_image(G, G.shape[1])
        gW, gb = [None] * L, [None] * L
        for l in range(L - 1, -1, -1):
            N, K = We[l].shape
            A = a_rows[l]
            # ---- dW_l = G_l^T A_l, db_l = column sums of G_l
            dW, db = _empty(N, K, like=g0), _empty(N, like=g0)
            if R >= WN_TC_WGRAD_MIN_ROWS and ctx.a_imgs:
                # split-K tcgen05 GEMM straight from the chain's own images (csrc/tc_mlp.cu tl_wgrad_mn_kernel): `img` is
                # the image of G_l (this layer's backward input), a_imgs[l] the image of its forward input
                work = torch.empty(lib().robir_tl_wgrad_mn_workspace(R, N, K, sm_count()), dtype=torch.uint8,
                                   device=g0.device)
                check(lib().robir_tl_wgrad_mn(ptr(img), (N + 63) // 64, ptr(ctx.a_imgs[l]), (K + 63) // 64, R, N, K,
                                              ptr(n_active), ptr(work), ptr(dW), ptr(db), sm_count(), stream()))
            elif R >= WN_TC_WGRAD_MIN_ROWS:
                # split-K tcgen05 GEMM over transposed hi/lo images (csrc/tc_mlp.cu tl_wgrad_kernel)
                work = torch.empty(lib().robir_tl_wgrad_workspace(R, N, K, sm_count()), dtype=torch.uint8, device=g0.device)
                check(lib().robir_tl_wgrad(ptr(G), G.shape[1], ptr(A), A.shape[1], R, N, K, ptr(n_active), ptr(work),
                                           ptr(dW), ptr(db), sm_count(), stream()))
            else:
                wt = ((N + 63) // 64) * ((K + 63) // 64)
                splits = max(1, min(64, (4 * sm_count()) // wt, (R + 63) // 64))
                part = _empty(splits * wt * 4160, like=g0) if splits > 1 else None
                tickets = _zeros(wt, dtype=torch.int32, like=g0)
                check(lib().robir_mlp_wgrad(ptr(G), G.shape[1], ptr(A), A.shape[1], R, N, K, None, 0, splits, ptr(part),
                                            ptr(tickets), ptr(dW), ptr(db), stream()))
            gW[l] = dW * (1.0 / math.sqrt(2.0)) if l in skip else dW
            gb[l] = db
            if l == 0:
                break
            # ---- G_{l-1} = (G_l W_l)[:, :N_{l-1}] * softplus'(h_{l-1}); h_{l-1} = the first N_{l-1} columns of A_l
            Np = K - d_in if l in skip else K
            # the fp32 copy of the hidden gradient is only read by the fp32-row weight-gradient kernels
            Gp = None if (R >= WN_TC_WGRAD_MIN_ROWS and ctx.a_imgs) else _empty(R, Np, like=g0)
            nkb_out = (Np + 63) // 64
            nxt = _tl_image(tiles, nkb_out, g0)
            q = _tl_params(img, ctx.imgs[l][1], None, R, Np, (N + 63) // 64, 1, SOFTPLUS, A, Gp, nxt, nkb_out, n_active, 0)
            q.no_fill = int(n_active is not None and R >= max(TL_BIG_MIN_ROWS, WN_TC_WGRAD_MIN_ROWS))
            _tl_layer(q, R)
            G, img = Gp, nxt
        return (None, None, None, None, *gW, *gb)


def wn_chain(x, Ws, bs, skip=(4,), n_active=None):
    """x [R, d_in] (no gradient), Ws[l] [N_l, K_l] the folded weight-norm weights, bs[l] [N_l] -> [R, N_last].
    n_active: optional device int32 [1], the number of leading rows that carry data (fixed-capacity batches, CUDA-graph
    mode): row tiles entirely beyond it are neither computed nor differentiated (outputs 0); the caller must feed zero
    upstream gradients for rows >= n_active."""
    return _WnChain.apply(x, tuple(skip), len(Ws), n_active, *Ws, *bs)


class _FusedMLP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, chain, x, extra, noise, noise_scale, want_param_grad, segments, *params):
        packed = chain.packed()
        need_bwd = want_param_grad or (extra is not None and extra.requires_grad) or x.requires_grad
        ctx.extra_shape = tuple(extra.shape) if extra is not None else None
        x = f32(x)
        n = x.shape[0]
        extra = f32(extra).reshape(-1) if extra is not None else None
        noise = f32(noise) if noise is not None else None
        n_out = packed[-1]["N"]
        out = _empty(n, n_out, like=x)
        n_active = active_rows.current
        p = _mlp_params(chain, packed, n, x, extra, noise, noise_scale, n_active, segments)
        saves, x0 = [], None
        if need_bwd:
            for l, d in enumerate(packed[:-1]):
                sv = _empty(n, d["Npad"], like=x)
                saves.append(sv)
                p.L[l].save = sv.data_ptr()
            if want_param_grad:
                x0 = _empty(n, packed[0]["Kpad"], like=x)
                p.x0_save = x0.data_ptr()
        p.out, p.ldo = ptr(out), n_out
        ctx.tc = chain.tc_ok(n)
        if ctx.tc:
            if x0 is None:
                x0 = _empty(n, packed[0]["Kpad"], like=x)
                p.x0_save = x0.data_ptr()
            ctx.tc_images = chain.packed_tc()      # shared with this step's backward, like ctx.packed
            _tl_forward(chain, packed, ctx.tc_images, p, n, x0, saves, out, n_active, segments)
        else:
            check(lib().robir_mlp_fwd(ctypes.byref(p), sm_count(), stream()))
        ctx.chain, ctx.n, ctx.want_param_grad = chain, n, want_param_grad
        ctx.packed = packed                 # the backward of this step uses the very same packed copies
        ctx.has_extra = extra is not None
        ctx.save_for_backward(x, extra, noise, x0, n_active, *saves)
        ctx.noise_scale, ctx.segments = noise_scale, segments
        return out

    @staticmethod
    def backward(ctx, g_out):
        chain, n = ctx.chain, ctx.n
        x, extra, noise, x0, n_active, *saves = ctx.saved_tensors
        packed = ctx.packed
        g_out = f32(g_out)
        p = _mlp_params(chain, packed, n, x, extra, noise, ctx.noise_scale, n_active, ctx.segments)
        for l, sv in enumerate(saves):
            p.L[l].save = sv.data_ptr()
        Gs = []
        if ctx.want_param_grad:
            for l, d in enumerate(packed):
                G = _zeros(n, d["Npad"], like=x)
                Gs.append(G)
                p.L[l].G = G.data_ptr()
        g_x = _empty(n, packed[0]["Kpad"], like=x)
        p.g_out, p.ldo, p.g_x = ptr(g_out), packed[-1]["N"], ptr(g_x)
        if ctx.tc:
            _tl_backward(chain, packed, ctx.tc_images, n, g_out, saves, Gs, g_x, n_active, ctx.segments,
                         ctx.want_param_grad)
        else:
            check(lib().robir_mlp_bwd(ctypes.byref(p), sm_count(), stream()))
        grads = []
        if ctx.want_param_grad:
            # dW_l = G_l^T A_{l-1} and db_l (csrc/mlp.cu wgrad_kernel): independent per layer -> parallel branches
            prevs = [x0] + list(saves)

            def layer_grads(l):
                d = packed[l]
                gw, gb = _empty(d["N"], d["K"], like=x), _empty(d["N"], like=x)
                seg = n // ctx.segments
                if n >= WN_TC_WGRAD_MIN_ROWS and ENGINE["mlp"] == "tc":
                    # large batches (the Vis stage trains VisNetwork on n_hit x 512 rows): split-K tcgen05 GEMM
                    work = torch.empty(lib().robir_tl_wgrad_workspace(n, d["N"], d["K"], sm_count()), dtype=torch.uint8,
                                       device=x.device)
                    check(lib().robir_tl_wgrad(ptr(Gs[l]), Gs[l].shape[1], ptr(prevs[l]), prevs[l].shape[1], n, d["N"],
                                               d["K"], ptr(n_active) if ctx.segments == 1 else None, ptr(work), ptr(gw),
                                               ptr(gb), sm_count(), stream()))
                    return gw, gb
                tiles = ((d["N"] + 63) // 64) * ((d["K"] + 63) // 64)
                splits = max(1, min(32, 160 // tiles, (n + 63) // 64))      # ~one wave of CTAs per layer
                part = _empty(splits * tiles * 4160, like=x) if splits > 1 else None
                check(lib().robir_mlp_wgrad(ptr(Gs[l]), Gs[l].shape[1], ptr(prevs[l]), prevs[l].shape[1], n, d["N"],
                                            d["K"], ptr(n_active) if seg % 32 == 0 else None, seg, splits, ptr(part),
                                            ptr(chain.wgrad_tickets(l, tiles, x)), ptr(gw), ptr(gb), stream()))
                return gw, gb
            if chain.wgrad_streams is None:
                chain.wgrad_streams = [torch.cuda.Stream() for _ in range(len(packed) - 1)]
            for gw, gb in fork_join([(lambda l=l: layer_grads(l)) for l in range(len(packed))], chain.wgrad_streams, tag="wgrad"):
                grads.append(gw)
                grads.append(gb)
        gx = g_x[:, :packed[0]["K"]] if chain.in_mode == 0 else None
        g_extra = g_x[:, 63].reshape(ctx.extra_shape) if ctx.has_extra else None
        return (None, gx, g_extra, None, None, None, None, *grads)


def fused_mlp(chain, x, extra=None, noise=None, noise_scale=0.02, want_param_grad=False, segments=1):
    """x: [n,K] (raw mode) or points [n,3]; extra: [n,1] appended column (pe10_extra); noise: [n,K] in embedding space;
    segments: the batch is a concatenation of that many equally ordered copies (matters under ops.active_rows)."""
    if extra is None and noise is None and not want_param_grad and active_rows.current is None and \
            chain.eval_tc_ok(x.shape[0], x):
        return chain.eval_tc(x)
    params = chain.params() if want_param_grad else []
    return _FusedMLP.apply(chain, x, extra, noise, noise_scale, want_param_grad, segments, *params)


# ----------------------------------------------------------------------------------------------------------------------
# fork / join of independent sub-graphs on side streams (captured as parallel branches of the step's CUDA graph)
# ----------------------------------------------------------------------------------------------------------------------
_side_streams = []


def _tensors_in(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _tensors_in(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _tensors_in(v)


_fork_state = {"depth": 0, "next": 0}
# debugging aids: ROBIR_FORK_SERIAL=1 runs every branch on the caller's stream, ROBIR_FORK_SERIAL_TAGS=nets,wgrad,...
# only the call sites with those tags (bisecting a suspected cross-stream hazard)
import os as _os
_FORK_SERIAL = _os.environ.get("ROBIR_FORK_SERIAL", "0") == "1"
_FORK_SERIAL_TAGS = set(_os.environ.get("ROBIR_FORK_SERIAL_TAGS", "").split(",")) - {""}


def fork_join(fns, streams=None, tag=""):
    """Run the callables concurrently: fns[0] on the current stream, the others on side streams that fork from and join
    back into it.  Host-side call order (hence the order of random draws) stays sequential.  Calls nest: every branch
    of one outermost call gets its own pooled stream, so sibling sub-branches never queue behind each other.
    streams: caller-owned side streams instead of the pool (autograd backward nodes, whose joins must not wait on
    unrelated work queued on the forward's pool streams)."""
    if _FORK_SERIAL or tag in _FORK_SERIAL_TAGS:
        return [fn() for fn in fns]
    main = torch.cuda.current_stream()
    st = _fork_state
    base = st["next"]
    if streams is None:
        st["next"] = base + len(fns) - 1
        while len(_side_streams) < st["next"]:
            _side_streams.append(torch.cuda.Stream(priority=-1))
        sides = _side_streams[base:base + len(fns) - 1]
    else:
        sides = list(streams)[:len(fns) - 1]
    st["depth"] += 1
    results = [None] * len(fns)
    try:
        fork = torch.cuda.Event()
        fork.record(main)
        for i, fn in enumerate(fns):          # host order = list order (keeps the random-draw order of the reference)
            if i == 0:
                results[0] = fn()
                continue
            side = sides[i - 1]
            side.wait_event(fork)
            with torch.cuda.stream(side):
                results[i] = fn()
        for side in sides:
            main.wait_stream(side)
    finally:
        st["depth"] -= 1
        if st["depth"] == 0:
            st["next"] = 0
    if not torch.cuda.is_current_stream_capturing():
        for r in results[1:]:
            for t in _tensors_in(r):
                t.record_stream(main)
    return results
