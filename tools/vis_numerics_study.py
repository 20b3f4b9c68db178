"""CPU emulation study of the number formats of the visibility-MLP tensor-core engine (numerics):
forward value and direction gradient of the VisNetwork (implicit_differentiable_renderer.py:225-258) on n (point,
direction) queries in fp64 (truth), fp32, bf16 hi/lo x3 (round 1), fp16 hi/lo x3 without / with power-of-two scaling
(round 2), single fp16 / bf16.  ReLU-unit flips are counted against fp64 and the gradient is re-measured with the fp64
masks forced, which separates "operand precision" from "a borderline unit took the other branch".

    python tools/vis_numerics_study.py [n]

Every matrix product of the hidden layers is evaluated as the engine evaluates it: operands rounded to the format's
hi (+ lo) parts, products summed exactly, result rounded to fp32; layer 0 and the 256 -> 2 head stay exact like the
engine's fp32 tables / FFMA head.  ``tests/test_numerics_study.py`` pins the conclusions on a smaller n."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from robir_b200 import synthetic  # noqa: E402

SA, SW, SG = 16.0, 64.0, 256.0          # activation / weight / unit-gradient scales of csrc/vis_tc.cu


def _pe(x, n_freq=10):
    out = [x]
    for l in range(n_freq):
        out += [torch.sin(x * 2.0 ** l), torch.cos(x * 2.0 ** l)]
    return torch.cat(out, -1)


def _round(x, fmt):
    return x.to(torch.bfloat16 if fmt == "bf16" else torch.float16).to(x.dtype)


def _split(x, fmt, scale):
    xs = x * scale
    hi = _round(xs.float(), fmt)
    lo = _round(xs.float() - hi, fmt)
    return hi.double() / scale, lo.double() / scale


def _mm(a, w, mode, sa, sw):
    fmt, terms = mode
    ah, al = _split(a, fmt, sa)
    wh, wl = _split(w, fmt, sw)
    r = ah @ wh.t()
    if terms == 3:
        r = r + al @ wh.t() + ah @ wl.t()
    return r.float()


def _engine_linear(w, mode, sa, sw, sg, dt):
    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a):
            return _mm(a, w, mode, sa, sw).to(dt)

        @staticmethod
        def backward(ctx, g):
            return _mm(g, w.t(), mode, sg, sw).to(dt)
    return F


def run(layers, p, d, gup, mode, dt, sa=1.0, sw=1.0, sg=1.0, masks=None):
    """-> (visibility [n] f64, d sum(vis * gup) / d dir [n,3] f64, ReLU masks of the four hidden layers)"""
    d = d.clone().to(dt).requires_grad_(True)
    h = torch.cat([_pe(p.to(dt)), _pe(d)], -1) @ layers[0][0].to(dt).t() + layers[0][1].to(dt)
    used = []
    for i in range(4):
        m = (h >= 0) if masks is None else masks[i]
        used.append(m.detach())
        h = torch.where(m, h, torch.zeros_like(h))
        if i == 3:
            break
        w, b = layers[i + 1]
        if mode in ("f64", "f32"):
            h = h @ w.to(dt).t() + b.to(dt)
        else:
            h = _engine_linear(w, mode, sa, sw, sg, dt).apply(h) + b.to(dt)
    z = h @ layers[4][0].to(dt).t() + layers[4][1].to(dt)
    vis = torch.softmax(z, -1)[:, 1]
    (vis * gup.to(dt)).sum().backward()
    return vis.detach().double(), d.grad.double(), used


def study(n=20000, seed=3):
    sd = synthetic.synthetic_state_dict(0, num_lgt_sgs=16)
    layers = [(sd["visibility_network.vis_layer.%d.weight" % i], sd["visibility_network.vis_layer.%d.bias" % i])
              for i in (0, 2, 4, 6, 8)]
    gen = torch.Generator().manual_seed(seed)
    p = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1) * 0.6
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    gup = torch.randn(n, generator=gen)
    v64, g64, m64 = run(layers, p, d, gup, "f64", torch.float64)
    rows = {}
    cases = [("fp32", "f32", 1, 1, 1), ("bf16 hi/lo x3 (round 1)", ("bf16", 3), 1, 1, 1),
             ("fp16 hi/lo x3, unscaled", ("f16", 3), 1, 1, 1), ("fp16 hi/lo x3, scaled (engine tc)", ("f16", 3), SA, SW, SG),
             ("fp16 single pass (engine tc1)", ("f16", 1), SA, SW, SG), ("bf16 single pass", ("bf16", 1), 1, 1, 1)]
    for name, mode, sa, sw, sg in cases:
        v, g, m = run(layers, p, d, gup, mode, torch.float32, sa, sw, sg)
        flips = sum(int((a != b).sum()) for a, b in zip(m, m64))
        _, gm, _ = run(layers, p, d, gup, mode, torch.float32, sa, sw, sg, masks=m64)
        rows[name] = dict(vis_max_abs=float((v - v64).abs().max()), grad_rel_l2=float((g - g64).norm() / g64.norm()),
                          flips=flips, grad_rel_l2_fp64_masks=float((gm - g64).norm() / g64.norm()))
    return rows, 4 * n * 256


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rows, units = study(n)
    print("%d queries, %d hidden units; errors vs fp64" % (n, units))
    print("%-36s %12s %14s %8s %22s" % ("format", "vis max abs", "grad rel L2", "flips", "grad rel L2, fp64 masks"))
    for k, r in rows.items():
        print("%-36s %12.2e %14.2e %8d %22.2e" % (k, r["vis_max_abs"], r["grad_rel_l2"], r["flips"],
                                                r["grad_rel_l2_fp64_masks"]))


if __name__ == "__main__":
    main()
