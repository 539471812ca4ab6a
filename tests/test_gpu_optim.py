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


def _compute_ema(value, ema, coeff, itr):
    """utils.py:389-396 restated."""
    import math
    if itr == 0:
        return value
    if itr <= int(math.floor(1 / coeff)):
        return (value + itr * ema) / (itr + 1)
    return coeff * value + (1 - coeff) * ema


def test_training_statistics_match_the_reference_loop_body():
    """csrc/ema.cu against the host-side bookkeeping of main.py:325-393 (compute_EMA, utils.py:389-396) replayed with
    torch ops on the same gradients / scalars, across the three regimes of compute_EMA (itr = 0, the running-mean
    warm-up, the exponential tail; EMA_coeff = 0.25 so that 7 iterations reach the tail)."""
    import soc_matching_b200 as sb
    d, hd, hm = 10, [256, 128, 64], [128, 128]
    st = random_setting("double_well", d, seed=2)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(st, seeded_unet(d, hd, 1), seeded_mnet(d, hm, 2), gam, hd, hm, DEV)
    params = list(sde.nabla_V.parameters())
    coeff, coeff_w = 0.25, 0.4
    stats = sb.TrainingStatistics(params, normalization_const=0.7, ema_coeff=coeff, ema_weight_mean_coeff=coeff_w)
    g = torch.Generator(DEV).manual_seed(3)
    ema_grad = ema_gn = ema_loss = ema_wm = ema_ws = None
    nc = torch.tensor(0.7, device=DEV)
    for itr in range(7):
        for p in params:
            p.grad = torch.randn(p.shape, device=DEV, generator=g) * (1.0 + 0.3 * itr)
        loss, wm, ws = (torch.rand((), device=DEV, generator=g) + 0.1 for _ in range(3))
        stats.update(loss, wm, ws, itr)
        grads = [p.grad.detach() for p in params]
        gn = sum(torch.norm(x) ** 2 for x in grads)
        if itr == 0:
            ema_grad, ema_gn, ema_loss, ema_wm, ema_ws = [x.clone() for x in grads], gn, loss, wm, ws
        else:
            ema_grad = [_compute_ema(x, e, coeff, itr) for x, e in zip(grads, ema_grad)]
            ema_gn = _compute_ema(gn, ema_gn, coeff, itr)
            ema_loss, ema_wm, ema_ws = (_compute_ema(v, e, coeff, itr) for v, e in ((loss, ema_loss), (wm, ema_wm), (ws, ema_ws)))
        nc = _compute_ema(wm, nc, coeff_w, itr)
        want = [gn, ema_gn, sum(torch.norm(e) ** 2 for e in ema_grad), ema_loss, ema_wm, ema_ws, nc]
        for name, w in zip(sb.training.STAT_NAMES, want):
            got = float(stats.as_dict()[name])
            assert abs(got - float(w)) <= 2e-6 * abs(float(w)), (itr, name, got, float(w))
        for e_mine, e_ref in zip(stats.ema_grad, ema_grad):
            assert rel_l2(e_mine, e_ref) <= 1e-6


def test_trainer_step_runs_the_loop_body():
    """Trainer.step = loss / normalisation -> backward -> statistics -> FusedAdam.step + zero_grad (main.py:298-393):
    parameters move, gradients are cleared, the normalisation constant becomes the EMA of mean(w)."""
    import soc_matching_b200 as sb
    d, hd, hm = 10, [256, 128, 64], [128, 128]
    st = random_setting("double_well", d, seed=2)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(st, seeded_unet(d, hd, 1), seeded_mnet(d, hm, 2), gam, hd, hm, DEV)
    solver = sb.SOC_Solver(sde, torch.zeros(d, device=DEV), None, T=1.0, num_steps=50, lmbd=1.0, d=d, sigma=sde.sigma)
    # main.py:188-230: lr 1e-4 for the control network, its own rates for the M-network and gamma
    opt = sb.FusedAdam([{"params": list(sde.nabla_V.parameters())},
                        {"params": list(sde.M.sigmoid_layers.parameters()), "lr": 1e-4},
                        {"params": [sde.gamma], "lr": 1e-3}], lr=1e-4, eps=1e-8)
    trainer = sb.Trainer(solver, opt, "SOCM", 256, normalization_const=0.5)
    before = [p.detach().clone() for p in sde.nabla_V.parameters()]
    wms = []
    for itr in range(3):
        loss, wm, ws = trainer.step(itr)
        wms.append(float(wm))
        assert torch.isfinite(loss)
    assert all(not torch.equal(b, p.detach()) for b, p in zip(before, sde.nabla_V.parameters()))
    assert all(float(p.grad.abs().max()) == 0.0 for p in sde.nabla_V.parameters())
    nc = wms[0]
    for itr in (1, 2):
        nc = _compute_ema(wms[itr], nc, 0.002, itr)
    assert abs(float(trainer.statistics.normalization_const) - nc) <= 1e-5 * abs(nc)
    # the solver (with its cached device state) stays picklable, as the reference's checkpoint needs (main.py:445-471)
    import pickle
    clone = pickle.loads(pickle.dumps(solver))
    assert torch.equal(clone.neural_sde.nabla_V.down_0[0].weight, sde.nabla_V.down_0[0].weight)
    out = clone.loss(64, algorithm="SOCM_const_M")
    assert torch.isfinite(out[0])
