"""Host-side logic of the CESR extras (robir_b200/cesr.py) that needs no GPU: state-dict layout, schedule, the
library-GEMM form of the weight-normed chain against the oracle, and the loud refusal of CPU tensors."""
import pytest
import torch

import pipeline as P
import robir_oracle as O
from robir_b200 import RobirError, cesr, synthetic


def test_wnmlp_state_dict_layout_and_init():
    """Keys / shapes of the reference SDFNetwork(d_in, d_out, 512, 8, [4], 0) (train_cesr.py:106-110; weight_norm keeps
    weight_g [o,1], weight_v [o,i], bias [o]) and the multires = 0 geometric initialisation (neus_model.py:358-376)."""
    for d_in, d_out in ((191, 2), (63, 3)):
        net = cesr.WnMLP(d_in, d_out)
        sd = net.state_dict()
        assert len(sd) == 27
        dims = [d_in] + [512] * 8 + [d_out]
        for l in range(9):
            o = dims[l + 1] - d_in if l + 1 == 4 else dims[l + 1]
            assert sd["lin%d.weight_v" % l].shape == (o, dims[l])
            assert sd["lin%d.weight_g" % l].shape == (o, 1)
            assert sd["lin%d.bias" % l].shape == (o,)
            assert torch.allclose(sd["lin%d.weight_g" % l], sd["lin%d.weight_v" % l].norm(dim=1, keepdim=True))
        assert torch.all(sd["lin8.bias"] == -0.5) and torch.all(sd["lin3.bias"] == 0)
        assert abs(sd["lin8.weight_v"].mean().item() - (torch.pi ** 0.5) / 512 ** 0.5) < 1e-4
    sh, nr = synthetic.cesr_state_dicts(0)
    assert set(sh) == set(cesr.WnMLP(191, 2).state_dict()) and set(nr) == set(cesr.WnMLP(63, 3).state_dict())


def test_library_form_of_the_chain_matches_oracle():
    sh, nr = synthetic.cesr_state_dicts(0)
    gen = torch.Generator().manual_seed(2)
    for sd, d_in in ((sh, 191), (nr, 63)):
        net = cesr.WnMLP(d_in, 2 if d_in == 191 else 3)
        net.load_state_dict(sd)
        x = torch.randn(70, d_in, generator=gen) * 0.5
        lins, skip = cesr._layers(net)
        Ws = [l.folded() for l in lins]
        out = cesr._wn_rows_torch(Ws, [l.bias for l in lins], skip, x)
        ref = O.wn_mlp(sd, "", x, prefix_dot=False)
        assert (out - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


def test_product_path_refuses_cpu_tensors():
    net = cesr.WnMLP(63, 3)
    with pytest.raises(RobirError):
        net(torch.zeros(4, 63))


@pytest.mark.parametrize("explore_iter,proj_iter", [(1000, 0), (0, 1000), (300, 200)])
def test_schedule_matches_reference_restatement(explore_iter, proj_iter):
    hook = cesr.ClusteredAlbedoHook(None, shadow_net=object(), normal_net=object(), explore_iter=explore_iter,
                                    proj_iter=proj_iter)
    for it in (0, 1, 499, 500, 501, 600, 999, 1000, 1001, 1200, 1499, 1500, 2750):
        hook.cur_iter = it
        assert hook.prefit_option() == P.cesr_prefit_option(it, explore_iter, proj_iter), it


class _EmulatedEngine:
    """Stand-in for the four C-ABI entry points that ops._WnChain drives, with the documented semantics of
    csrc/tc_mlp.cu / csrc/mlp.cu restated in torch on CPU tensors (images are tracked as logical matrices).  It checks
    the HOST logic of the chain -- buffer plumbing, skip concat, gradient chain, scaling -- not the kernels."""

    def __init__(self):
        self.mats = {}

    @staticmethod
    def _act(x, act):
        return torch.nn.functional.softplus(x, beta=100) if act == 3 else x

    @staticmethod
    def _dact(y, act):
        return -torch.expm1(-100.0 * y) if act == 3 else torch.ones_like(y)

    def robir_tl_pack_weight(self, W, ldw, N, K, transpose, col_blocks, nkb, img, stream):
        assert W.shape[1] == ldw and col_blocks * 128 >= N and nkb * 64 >= K
        B = W.t()[:N, :K] if transpose else W[:N, :K]
        assert B.shape == (N, K)
        assert img.numel() == col_blocks * nkb * 32768
        self.mats[id(img)] = (B.clone(), nkb)
        return 0

    def robir_tl_pack_rows(self, X, ldx, n, K, ref, ld_ref, act, nkb, img, stream):
        assert X.shape[1] == ldx and nkb * 64 >= K and ref is None
        assert img.numel() == ((n + 127) // 128) * nkb * 32768
        self.mats[id(img)] = (X[:n, :K].clone(), nkb)
        return 0

    def robir_tl_layer(self, q, stream):
        (A, nkb_a), (B, nkb_b) = self.mats[id(q.a_img)], self.mats[id(q.w_img)]
        assert nkb_a == nkb_b == q.nkb <= 8 and A.shape[0] == q.n
        assert B.shape[0] == q.N, "weight image rows must equal the valid output columns"
        K = min(A.shape[1], B.shape[1])
        assert not A[:, K:].any() and not B[:, K:].any()        # anything beyond the shorter operand must be padding
        acc = A[:, :K] @ B[:, :K].t()
        if q.mode == 0:
            assert q.bias.numel() >= ((q.N + 127) // 128) * 128
            x = self._act(acc + q.bias[:q.N], q.act)
        else:
            x = acc * (self._dact(q.ref[:, :q.N], q.act) if q.ref is not None else 1.0)
        if q.out is not None:
            assert q.ld_out == q.out.shape[1]
            q.out[:, :q.N] = x
        if q.out_img is not None:
            assert q.nkb_out == (q.N + 63) // 64
            self.mats[id(q.out_img)] = (x.clone(), q.nkb_out)
        return 0

    def robir_mlp_wgrad(self, G, ldg, A, lda, n, N, K, n_active, seg, splits, partial, tickets, dW, db, stream):
        assert G.shape[1] == ldg and A.shape[1] == lda and 1 <= splits <= 64
        assert splits == 1 or partial.numel() >= splits * ((N + 63) // 64) * ((K + 63) // 64) * 4160
        dW.copy_(G[:n, :N].t() @ A[:n, :K])
        db.copy_(G[:n, :N].sum(0))
        return 0

    def robir_tl_wgrad_workspace(self, n, N, K, sm_count):
        return 1024

    def robir_tl_wgrad_mn_workspace(self, n, N, K, sm_count):
        return 1024

    def robir_tl_wgrad_mn(self, g_img, nkb_g, a_img, nkb_a, n, N, K, n_active, work, dW, db, sm_count, stream):
        (G, kg), (A, ka) = self.mats[id(g_img)], self.mats[id(a_img)]
        assert kg == nkb_g == (N + 63) // 64 and ka == nkb_a == (K + 63) // 64 and n_active is None
        assert G.shape == (n, N) and A.shape == (n, K) and work.numel() >= 1024
        self.tc_wgrads = getattr(self, "tc_wgrads", 0) + 1
        dW.copy_(G.t() @ A)
        db.copy_(G.sum(0))
        return 0

    def robir_tl_wgrad(self, G, ldg, A, lda, n, N, K, n_active, work, dW, db, sm_count, stream):
        assert G.shape[1] == ldg and A.shape[1] == lda and work.numel() >= 1024 and n_active is None
        self.tc_wgrads = getattr(self, "tc_wgrads", 0) + 1
        dW.copy_(G[:n, :N].t() @ A[:n, :K])
        db.copy_(G[:n, :N].sum(0))
        return 0


@pytest.mark.parametrize("d_in,d_out,rows,tc_wgrad", [(191, 2, 256, False), (191, 2, 200, True), (63, 3, 130, False)])
def test_wn_chain_host_logic_with_emulated_engine(monkeypatch, d_in, d_out, rows, tc_wgrad):
    import types
    from robir_b200 import ops
    eng = _EmulatedEngine()
    monkeypatch.setattr(ops, "WN_TC_WGRAD_MIN_ROWS", 128 if tc_wgrad else 1 << 30)

    def params(a_img, w_img, bias, n, N, nkb, mode, act, ref, out, out_img, nkb_out, n_active, seg):
        return types.SimpleNamespace(a_img=a_img, w_img=w_img, bias=bias, n=n, N=N, nkb=nkb, mode=mode, act=act, ref=ref,
                                     out=out, ld_out=out.shape[1] if out is not None else 0, out_img=out_img,
                                     nkb_out=nkb_out)
    monkeypatch.setattr(ops, "lib", lambda: eng)
    monkeypatch.setattr(ops, "ptr", lambda t: t)
    monkeypatch.setattr(ops, "stream", lambda: None)
    monkeypatch.setattr(ops, "check", lambda status: None)
    monkeypatch.setattr(ops, "sm_count", lambda: 148)
    monkeypatch.setattr(ops, "_tl_params", params)
    monkeypatch.setattr(ops, "ctypes", types.SimpleNamespace(byref=lambda q: q))

    sh, nr = synthetic.cesr_state_dicts(0)
    net = cesr.WnMLP(d_in, d_out)
    net.load_state_dict(sh if d_in == 191 else nr)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(rows, d_in, generator=gen) * 0.5
    gup = torch.randn(rows, d_out, generator=gen)
    lins, skip = cesr._layers(net)
    res = []
    for fn in (lambda Ws, bs: ops.wn_chain(x, Ws, bs, skip), lambda Ws, bs: cesr._wn_rows_torch(Ws, bs, skip, x)):
        net.zero_grad()
        out = fn([l.folded() for l in lins], [l.bias for l in lins])
        (out * gup).sum().backward()
        res.append((out.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()}))
    (o1, g1), (o2, g2) = res
    with torch.no_grad():                 # evaluation mode: same values, nothing kept for a backward
        o3 = ops.wn_chain(x, [l.folded() for l in lins], [l.bias for l in lins], skip)
    assert torch.equal(o3, o1) and not o3.requires_grad
    assert (o1 - o2).abs().max().item() < 1e-5 * max(1.0, o2.abs().max().item())
    assert set(g1) == set(g2) and len(g1) == 27
    assert getattr(eng, "tc_wgrads", 0) == (9 if tc_wgrad else 0)      # one launch chain per layer on the selected engine
    for k in g2:
        assert (g1[k] - g2[k]).abs().max().item() < 2e-4 * max(1e-6, g2[k].abs().max().item()), k


def test_hook_binds_to_a_live_runner_and_follows_its_counters():
    """ClusteredAlbedoHook.bind: conf keys of confs_sg/*.conf train{}, live cur_iter / is_training of the runner."""
    import types

    class Conf(dict):
        get_int = lambda self, k: int(self[k])
        get_float = lambda self, k: float(self[k])

    model = types.SimpleNamespace(get_sg_render=None)
    runner = types.SimpleNamespace(
        model=model, shadow_net=object(), normal_net=object(), white_light=True, cur_iter=0, is_training=True,
        train_spec=True, conf=Conf({'train.explore_iter': 0, 'train.proj_iter': 1000, 'train.explore_smooth': 0.001,
                                    'train.explore_kl': 0.01, 'train.proj_smooth': 0.002, 'train.proj_kl': 0.03}))
    hook = cesr.ClusteredAlbedoHook.bind(runner)
    assert model.get_sg_render == hook.get_sg_render and hook.white_light
    for it in (10, 700, 1500):
        runner.cur_iter = it
        assert hook.cur_iter == it and hook.prefit_option() == P.cesr_prefit_option(it, 0, 1000)
    runner.is_training = False
    assert hook.is_training is False
    assert hook.weights == dict(explore=(0.001, 0.01), project=(0.002, 0.03))
    # stand-alone use keeps its own counters
    own = cesr.ClusteredAlbedoHook(None, object(), object(), cur_iter=5)
    own.cur_iter = 900
    assert own.cur_iter == 900 and own.prefit_option() == "explore"


def test_fixed_capacity_forward_dispatches_on_the_bound_hook():
    """The fixed-capacity (CUDA-graph) forward inlines the PBR hook and knows the CESR hook's fixed-capacity form; any
    other re-bound get_sg_render must be refused instead of silently rendering the PBR stage."""
    import robir_b200
    model = robir_b200.IDRNetwork(dict(envmap_material_network=dict(num_lgt_sgs=128)))
    hook = cesr.ClusteredAlbedoHook(model, object(), object(), cur_iter=600)
    inp = {"intrinsics": None, "hdr_shift": None, "uv": None, "pose": None, "object_mask": None}
    model.static_shapes = True
    default = model.get_sg_render
    seen = {}
    model._forward_static = lambda input, lin_diff=False, train_spec=False, hook=None: seen.setdefault("hook", hook)
    model.get_sg_render = hook.get_sg_render
    model(inp, trainstage="Material", train_spec=True)
    assert seen.pop("hook") is hook                     # CESR: the hook's get_sg_render_static is what will run
    model.get_sg_render = default                       # restoring the default bound method is not a re-binding
    model(inp, trainstage="Material", train_spec=True)
    assert "hook" in seen and seen.pop("hook") is None
    model.get_sg_render = lambda *a, **k: {}            # anything else
    with pytest.raises(RobirError, match="static_shapes"):
        model(inp, trainstage="Material", train_spec=True)

    class Other:
        def get_sg_render(self, *a, **k):
            return {}
    model.get_sg_render = Other().get_sg_render
    with pytest.raises(RobirError, match="static_shapes"):
        model(inp, trainstage="Material", train_spec=True)


def test_phase_key_follows_the_schedule():
    """One CUDA graph per phase: the key changes exactly where the step's control flow does (train_cesr.py:546-559, :508)."""
    hook = cesr.ClusteredAlbedoHook(object(), object(), object(), cur_iter=0)
    keys = {}
    for it in (0, 300, 500, 501, 600, 1000, 1001, 1500):
        hook.cur_iter = it
        keys[it] = hook.phase_key()
    assert keys[0] == keys[300] == keys[500]
    assert keys[501] == keys[600] == keys[1000] != keys[500]
    assert keys[1001] == keys[1500] != keys[1000]
    hook.is_training = False
    assert hook.phase_key() != keys[1500]


def test_shadow_net_rank_structure_of_the_one_hot_input():
    """Design check for DESIGN.md section 6 item 2d: shadow_net's input rows are [PE(x_i) | onehot(m)], so layer 0 and the
    skip columns of layer 4 are sums of a per-point table and a per-lobe table -- no 191-wide GEMM over n*128 rows."""
    sh, _ = synthetic.cesr_state_dicts(0)
    net = cesr.WnMLP(191, 2)
    net.load_state_dict(sh)
    gen = torch.Generator().manual_seed(8)
    n, M = 5, 128
    emb = O.pe(torch.randn(n, 3, generator=gen) * 0.3, 10)
    x = torch.cat([emb[:, None, :].expand(-1, M, -1), torch.eye(M)[None].expand(n, -1, -1)], -1).reshape(n * M, -1)
    lins, skip = cesr._layers(net)
    Ws, bs = [l.folded().detach() for l in lins], [l.bias.detach() for l in lins]
    full = cesr._wn_rows_torch(Ws, bs, skip, x)
    sp = lambda t: torch.nn.functional.softplus(t, beta=100)
    h = sp((emb @ Ws[0][:, :63].t() + bs[0])[:, None, :] + Ws[0][:, 63:].t()[None, :, :]).reshape(n * M, -1)
    for l in (1, 2, 3):
        h = sp(h @ Ws[l].t() + bs[l])
    k = Ws[4].shape[1] - 191
    tab = ((emb @ Ws[4][:, k:k + 63].t())[:, None, :] + Ws[4][:, k + 63:].t()[None, :, :]).reshape(n * M, -1)
    h = sp((h @ Ws[4][:, :k].t() + tab) / 2 ** 0.5 + bs[4])
    for l in (5, 6, 7):
        h = sp(h @ Ws[l].t() + bs[l])
    out = h @ Ws[8].t() + bs[8]
    assert (out - full).abs().max().item() < 1e-5 * max(1.0, full.abs().max().item())
