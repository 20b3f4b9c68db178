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
