"""Tabulated ground-truth controls (models.py:10-150) in the rollout: ``control_objective(optimal_sde, ...)`` of
main.py:138 runs under LinearControl (OU_quadratic), ConstantControlLinear (OU_linear) or LowDimControl (double_well).
CPU: the mirror classes against the reference's own classes (when /root/reference is present).  GPU: the fused
tabulated rollout against the oracle's rollout driven by the same control (all 8 outputs, trajectories 1e-5)."""
import os
import sys

import pytest
import torch

from helpers import orc, random_setting, rel_l2

REF = "/root/reference"


def _controls(d, K, seed=0):
    import soc_matching_b200 as sb
    g = torch.Generator().manual_seed(seed)
    lin = sb.LinearControl(0.4 * torch.randn(K + 1, d, d, generator=g), 1.0)
    const = sb.ConstantControlLinear(torch.randn(K + 1, d, generator=g), 1.0)
    nt, nx = 2 * K + 1, 41
    low = sb.LowDimControl(torch.randn(nt, nx, d, generator=g), 1.0, 2.75, d, 1.0 / (2 * K), 5.5 / (nx - 1))
    return {"linear": lin, "constant": const, "lowdim": low}


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_mirror_classes_match_reference_classes():
    sys.path.insert(0, REF)
    from SOC_matching import models as rm
    d, K, B = 3, 12, 7
    ours = _controls(d, K)
    refs = {"linear": rm.LinearControl(ours["linear"].u, 1.0), "constant": rm.ConstantControlLinear(ours["constant"].ut, 1.0),
            "lowdim": rm.LowDimControl(ours["lowdim"].ut, 1.0, 2.75, d, ours["lowdim"].delta_t, ours["lowdim"].delta_x)}
    ts = torch.linspace(0, 1.0, K + 1)
    x2 = 1.5 * torch.randn(B, d, generator=torch.Generator().manual_seed(1))
    x3 = 1.5 * torch.randn(K + 1, B, d, generator=torch.Generator().manual_seed(2))
    for name in ours:
        for k in (0, 5, K - 1):
            # same table row, same product (matmul vs einsum: last-bit differences only for the linear control)
            assert torch.allclose(ours[name](ts[k], x2), refs[name](ts[k], x2), rtol=1e-6, atol=1e-7), (name, k)
        assert torch.allclose(ours[name](ts, x3, t_is_tensor=True), refs[name](ts, x3, t_is_tensor=True), rtol=1e-6,
                              atol=1e-7), name


@pytest.mark.gpu
@pytest.mark.parametrize("kind,ctrl", [("ou_quadratic", "linear"), ("ou_linear", "constant"), ("double_well", "lowdim"),
                                        ("molecular_dynamics", "lowdim")])
def test_tabulated_rollout_matches_oracle(kind, ctrl):
    import soc_matching_b200 as sb
    from helpers import make_product_sde, seeded_mnet, seeded_unet
    DEV = "cuda"
    d = 1 if kind == "molecular_dynamics" else 6
    K, B = 40, 77
    st = random_setting(kind, d, seed=5, dense_sigma=(kind == "ou_linear"))
    u = _controls(d, K, seed=3)[ctrl]
    ts = torch.linspace(0, 1.0, K + 1)
    x0 = -torch.ones(d) if kind == "molecular_dynamics" else 0.2 * torch.ones(d)
    noises = torch.randn(K, B, d, generator=torch.Generator().manual_seed(9))
    want = orc.rollout(st, None, x0.repeat(B, 1), ts, noises=noises, control_fn=lambda k, t0, x: u(t0, x))
    gam = {"gamma": torch.tensor([2.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(st, seeded_unet(d, [16, 8, 8], 1), seeded_mnet(d, [8, 8], 2), gam, [16, 8, 8], [8, 8], DEV)
    sde.use_learned_control = False                      # the optimal SDE of settings.py:25-114: u = tabulated control
    moved = {"linear": lambda: sb.LinearControl(u.u.to(DEV), 1.0), "constant": lambda: sb.ConstantControlLinear(u.ut.to(DEV), 1.0),
             "lowdim": lambda: sb.LowDimControl(u.ut.to(DEV), 1.0, u.xb, d, u.delta_t, u.delta_x)}[ctrl]()
    sde.u = moved
    got = sb.stochastic_trajectories(sde, x0.to(DEV).repeat(B, 1), ts.to(DEV), st.lmbd, noises=noises.to(DEV))
    names = ["states", "noises", "stop_indicators", "fractional_timesteps", "logw_det", "logw_sto", "logw_term", "controls"]
    for key, a, b in zip(names, got, want):
        a, b = a.detach().float().cpu(), b.detach().float().cpu()
        if key == "stop_indicators":
            assert torch.equal(a, b), key
        else:
            assert rel_l2(a, b) <= (1e-4 if key.startswith("logw") else 1e-5), (key, rel_l2(a, b))
    # control_objective in weights-only mode (utils.py:131-163): mean of -lmbd (logw_det + logw_term) over fresh paths
    mean, err = sb.control_objective(sde, x0.to(DEV), ts.to(DEV), st.lmbd, 64, total_n_samples=4096)
    assert torch.isfinite(mean) and torch.isfinite(err) and float(err) > 0
