"""FusedAdam (csrc/adam.cu) against torch.optim.Adam with the reference's parameter groups (main.py:188-230:
UNet, M-network and gamma with their own learning rates): parameters and both moment buffers after several steps
agree to 1e-6, gradients are cleared when asked."""
import pytest
import torch

from helpers import make_product_sde, random_setting, rel_l2, seeded_mnet, seeded_unet

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _groups(sde, y0):
    return [{"params": list(sde.nabla_V.parameters())},
            {"params": list(sde.M.sigmoid_layers.parameters()), "lr": 3e-3},
            {"params": [sde.gamma], "lr": 1e-2},
            {"params": [y0], "lr": 5e-2}]


def test_fused_adam_matches_torch_adam():
    import soc_matching_b200 as sb
    d, hd, hm = 10, [256, 128, 64], [128, 128]
    st = random_setting("double_well", d, seed=2)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sdes = [make_product_sde(st, seeded_unet(d, hd, 1), seeded_mnet(d, hm, 2), gam, hd, hm, DEV) for _ in range(2)]
    y0s = [torch.nn.Parameter(torch.tensor([0.3], device=DEV)) for _ in range(2)]
    ref = torch.optim.Adam(_groups(sdes[0], y0s[0]), lr=1e-3, eps=1e-8)
    mine = sb.FusedAdam(_groups(sdes[1], y0s[1]), lr=1e-3, eps=1e-8)
    g = torch.Generator(DEV).manual_seed(0)
    for it in range(5):
        for (pa, pb) in zip([p for grp in ref.param_groups for p in grp["params"]],
                            [p for grp in mine.param_groups for p in grp["params"]]):
            grad = torch.randn(pa.shape, device=DEV, generator=g) * (0.1 + it)
            pa.grad, pb.grad = grad.clone(), grad.clone()
        ref.step()
        mine.step(zero_grad=(it == 4))
    pa_all = [p for grp in ref.param_groups for p in grp["params"]]
    pb_all = [p for grp in mine.param_groups for p in grp["params"]]
    for pa, pb in zip(pa_all, pb_all):
        assert rel_l2(pb.detach(), pa.detach()) <= 1e-6
        assert rel_l2(mine.state[pb]["exp_avg"], ref.state[pa]["exp_avg"]) <= 1e-6
        assert rel_l2(mine.state[pb]["exp_avg_sq"], ref.state[pa]["exp_avg_sq"]) <= 1e-6
        assert float(pb.grad.abs().max()) == 0.0
    # a training iteration end to end: loss -> backward -> fused step changes every parameter group
    sde = sdes[1]
    solver = sb.SOC_Solver(sde, torch.zeros(d, device=DEV), None, T=1.0, num_steps=20, lmbd=1.0, d=d, sigma=sde.sigma)
    before = [p.detach().clone() for p in pb_all[:-1]]
    out = solver.loss(256, algorithm="SOCM")
    (out[0] / out[5].detach()).backward()
    mine.step(zero_grad=True)
    assert all(not torch.equal(b, p.detach()) for b, p in zip(before, pb_all[:-1]))
