"""GPU parity of the SOCM iteration (K1 -> prep -> K2 -> K3 -> K2^T -> M-network autograd) through
the public drop-in ``SOC_Solver.loss`` + ``backward``: against the reference's own loss and
gradients (tests/golden) and against the oracle at the default network width.
Tolerance (north star): loss and per-tensor gradients within 1e-4 relative (norm-wise)."""
import pytest
import torch

from helpers import (Golden, golden_names, make_product_sde, orc, random_setting, rel_l2, seeded_mnet,
                     seeded_unet)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run_product(sde, x0, K, B, lmbd, noises, algo, stopping=False, warm=None, generic=False, chunk=None,
                ffma=False, y0=None, add_weights=False):
    import soc_matching_b200 as sb
    solver = sb.SOC_Solver(sde, x0.to(DEV), None, T=1.0, num_steps=K, lmbd=lmbd, d=sde.dim, sigma=sde.sigma)
    if y0 is not None:
        solver.y0.data.copy_(y0.to(DEV))
    solver.y0.grad = None
    if algo == "SOCM_exp":   # main.py:166-169
        solver.gamma = torch.nn.Parameter(sde.gamma.detach().clone())
    solver.force_generic = generic
    solver.force_ffma = ffma
    if chunk:
        solver.chunk_paths = chunk
    for p in sde.parameters():
        p.grad = None
    solver.inject_noise(noises.to(DEV))
    out = solver.loss(B, algorithm=algo, u_warm_start=sde.u_warm_start if warm is not None else None,
                      use_warm_start=warm is not None, use_stopping_time=stopping, add_weights=add_weights)
    out[0].backward()
    grads = {} if solver.y0.grad is None else {"gam/y0": solver.y0.grad}
    for n, p in sde.nabla_V.named_parameters():
        grads["unet/" + n] = p.grad
    for n, p in sde.M.sigmoid_layers.named_parameters():
        grads["mnet/sigmoid_layers." + n] = p.grad
    grads["gam/gamma"] = solver.gamma.grad if algo == "SOCM_exp" else sde.gamma.grad
    if stopping:
        grads["gam/gamma2"], grads["gam/gamma3"] = sde.gamma2.grad, sde.gamma3.grad
    return out, grads


def check_grads(got, want, tol=1e-4, gamma_tol=None):
    for key, w in want.items():
        g = got.get(key)
        g = torch.zeros_like(w) if g is None else g.detach().cpu()
        t = gamma_tol if (gamma_tol and key.startswith("gam/")) else tol
        assert rel_l2(g, w) <= t, (key, rel_l2(g, w))


@pytest.mark.parametrize("name", golden_names())
def test_loss_and_grads_match_reference_golden(name):
    g = Golden(name)
    m = g.meta
    for algo in m["algorithms"]:
        sde = make_product_sde(g.setting, g.unet, g.mnet, g.gammas, m["hdims"], m["hdims_M"], DEV,
                               stopping=m["stopping"], warm=g.warm)
        out, grads = run_product(sde, g.x0, m["K"], m["B"], m["lmbd"], g.traj[1], algo, stopping=m["stopping"],
                                 warm=g.warm, y0=g.gammas.get("y0"))
        if m["hdims"] == [256, 128, 64]:   # default width runs the tcgen05 kernels: check the FFMA tile path too
            sde2 = make_product_sde(g.setting, g.unet, g.mnet, g.gammas, m["hdims"], m["hdims_M"], DEV,
                                    stopping=m["stopping"], warm=g.warm)
            out2, grads2 = run_product(sde2, g.x0, m["K"], m["B"], m["lmbd"], g.traj[1], algo,
                                       stopping=m["stopping"], warm=g.warm, ffma=True, y0=g.gammas.get("y0"))
            assert abs(float(out2[0]) - g.scalar(f"{algo}/loss")) <= 1e-4 * abs(g.scalar(f"{algo}/loss"))
            check_grads(grads2, g.grads(algo), gamma_tol=2e-3 if m["stopping"] else None)
        want = g.scalar(f"{algo}/loss")
        assert abs(float(out[0]) - want) <= 1e-4 * abs(want), (algo, float(out[0]), want)
        assert abs(float(out[5]) - g.scalar(f"{algo}/weight_mean")) <= 1e-4 * abs(g.scalar(f"{algo}/weight_mean"))
        assert abs(float(out[6]) - g.scalar(f"{algo}/weight_std")) <= 1e-3 * abs(g.scalar(f"{algo}/weight_std")) + 1e-9
        assert torch.equal(out[7].cpu(), g.traj[2])
        # d/dgamma of the stopping-time M is ill-conditioned (2e-4 between AD modes)
        check_grads(grads, g.grads(algo), gamma_tol=2e-3 if m["stopping"] else None)


CASES = [  # kind, d, K, B, dense sigma
    ("double_well", 10, 60, 70, False),
    ("ou_quadratic", 20, 12, 40, False),
    ("ou_linear", 10, 16, 33, True),
]


@pytest.mark.parametrize("kind,d,K,B,dense", CASES)
@pytest.mark.parametrize("algo", ["SOCM", "SOCM_const_M", "SOCM_exp", "SOCM_adjoint", "cross_entropy", "log-variance"])
def test_default_width_matches_oracle(kind, d, K, B, dense, algo):
    st = random_setting(kind, d, seed=d + K, dense_sigma=dense)
    hd, hm = [256, 128, 64], [128, 128]
    unet = seeded_unet(d, hd, 21 + d)
    mnet = seeded_mnet(d, hm, 22 + d, 0.1)
    gam = {"gamma": torch.tensor([2.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    x0 = torch.zeros(d) if kind == "double_well" else 0.3 * torch.ones(d)
    ts = torch.linspace(0, 1.0, K + 1)
    noises = torch.randn(K, B, d, generator=torch.Generator().manual_seed(5))
    torch.set_num_threads(8)
    traj = orc.rollout(st, unet, x0.repeat(B, 1), ts, noises=noises)
    pu = {k: v.clone().requires_grad_(True) for k, v in unet.items()}
    pm = {k: v.clone().requires_grad_(True) for k, v in mnet.items()}
    pg = {k: v.clone().requires_grad_(True) for k, v in gam.items()}
    obj, wm, wsd = orc.socm_loss(st, pu, pm, pg, ts, traj, algorithm=algo)
    obj.backward()
    want = {"unet/" + k: v.grad for k, v in pu.items()}
    if algo == "SOCM":
        want.update({"mnet/" + k: v.grad for k, v in pm.items()})
    if algo in ("SOCM", "SOCM_exp"):
        want["gam/gamma"] = pg["gamma"].grad
    # default dispatch (fp32 FFMA tile K3 below SOCM_LOSS_TC_MIN_POINTS; tcgen05 rollout) and shape-generic;
    # the tcgen05 K3 is covered by tests/test_gpu_loss_tc.py
    for kernel in ("auto", "ffma", "generic"):
        sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
        out, grads = run_product(sde, x0, K, B, st.lmbd, noises, algo, generic=(kernel == "generic"),
                                 ffma=(kernel == "ffma"))
        # cross_entropy = mean(S_m w_m) with S_m of both signs: the value can cancel to ~0, so it is compared on the
        # scale of its terms (mean weight) rather than relative to itself
        base = max(abs(float(obj)), float(wm)) if algo == "cross_entropy" else abs(float(obj))
        assert abs(float(out[0]) - float(obj)) <= 1e-4 * base, (kernel, float(out[0]), float(obj))
        assert abs(float(out[5]) - float(wm)) <= 1e-4 * abs(float(wm))
        check_grads(grads, want)


def test_chunked_equals_unchunked_and_scales_with_grad_output():
    st = random_setting("double_well", 10, seed=1)
    hd, hm = [256, 128, 64], [128, 128]
    unet, mnet = seeded_unet(10, hd, 3), seeded_mnet(10, hm, 4)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    K, B = 60, 150
    noises = torch.randn(K, B, 10, generator=torch.Generator().manual_seed(8))
    res = []
    for chunk in (None, 64):
        sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
        res.append(run_product(sde, torch.zeros(10), K, B, 1.0, noises, "SOCM", chunk=chunk))
    (o1, g1), (o2, g2) = res
    assert torch.isfinite(o1[0]) and abs(float(o1[0]) - float(o2[0])) <= 1e-5 * abs(float(o1[0]))
    assert o1[7].shape == o2[7].shape == (K + 1, B)
    for k in g1:
        assert rel_l2(g2[k], g1[k]) <= 2e-5, (k, rel_l2(g2[k], g1[k]))
    # (loss / c).backward() as in main.py:320-323
    import soc_matching_b200 as sb
    sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
    solver = sb.SOC_Solver(sde, torch.zeros(10, device=DEV), None, num_steps=K, lmbd=1.0, d=10, sigma=sde.sigma)
    solver.inject_noise(noises.to(DEV))
    out = solver.loss(B, algorithm="SOCM")
    (out[0] / 4.0).backward()
    for n, p in sde.nabla_V.named_parameters():
        assert rel_l2(p.grad * 4.0, g1["unet/" + n]) <= 1e-6
    assert rel_l2(sde.gamma.grad * 4.0, g1["gam/gamma"]) <= 1e-5


def test_unsupported_requests_raise():
    import soc_matching_b200 as sb
    st = random_setting("double_well", 4, seed=1)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(st, seeded_unet(4, [16, 8, 8], 1), seeded_mnet(4, [8, 8], 2), gam, [16, 8, 8], [8, 8], DEV)
    solver = sb.SOC_Solver(sde, torch.zeros(4, device=DEV), None, num_steps=5, lmbd=1.0, d=4, sigma=sde.sigma)
    with pytest.raises(NotImplementedError):
        solver.loss(8, algorithm="rel_entropy")
    with pytest.raises(sb._lib.SocmError):
        solver.loss(8, algorithm="SOCM", use_stopping_time=True)


def test_l2_error_and_normalization_constant():
    """Evaluation-side rows of SURVEY.md section 8f: the importance-weighted L2 error of ``loss`` (method.py:858-875)
    and ``normalization_constant`` (utils.py:166-231) against the same formulas evaluated with torch on the
    trajectories the kernels produced (the rollout itself is pinned by tests/test_gpu_rollout.py)."""
    import math
    from types import SimpleNamespace
    import soc_matching_b200 as sb
    from soc_matching_b200 import simulate
    d, K, B = 4, 20, 48
    st = random_setting("ou_quadratic", d, seed=11)
    hd, hm = [24, 16, 8], [12, 12]
    unet, mnet = seeded_unet(d, hd, 3), seeded_mnet(d, hm, 4, 0.1)
    gam = {"gamma": torch.tensor([2.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
    x0 = 0.3 * torch.ones(d, device=DEV)
    Kmat = torch.randn(d, d, generator=torch.Generator().manual_seed(1)).to(DEV) * 0.3

    def optimal_control(ts, states, t_is_tensor=True):       # any torch callable with the reference's signature
        return -torch.einsum("ij,abj->abi", Kmat, states) * (1.0 - ts.reshape(-1, 1, 1))

    solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=K, lmbd=st.lmbd, d=d, sigma=sde.sigma)
    noises = torch.randn(K, B, d, generator=torch.Generator().manual_seed(5)).to(DEV)
    solver.inject_noise(noises)
    out = solver.loss(B, compute_L2_error=True, optimal_control=optimal_control, algorithm="SOCM_adjoint")
    ts = torch.linspace(0, 1.0, K + 1)
    traj = orc.rollout(st, unet, x0.cpu().repeat(B, 1), ts, noises=noises.cpu())
    gv = orc.nabla_v_all(st, unet, ts, traj[0], None)
    learned = -torch.einsum("ij,abj->abi", st.sigma.t(), gv)
    w = torch.exp(traj[4] + traj[5] + traj[6])
    want = torch.sum((optimal_control(ts, traj[0]).cpu() if False else
                      (-torch.einsum("ij,abj->abi", Kmat.cpu(), traj[0]) * (1.0 - ts.reshape(-1, 1, 1))) - learned) ** 2
                     * w.reshape(1, -1, 1)) / ((K + 1) * B)
    assert abs(float(out[1]) - float(want)) <= 1e-4 * abs(float(want)), (float(out[1]), float(want))

    # normalization_constant: replay its rollouts with the same Philox keys
    cfg = SimpleNamespace(method=SimpleNamespace(lmbd=st.lmbd))
    xb = x0.repeat(B, 1)
    c0 = simulate._SEED_COUNTER[0]
    nc, nc_err, sqd = sb.normalization_constant(sde, xb, solver.ts, cfg, n_batches_normalization=6,
                                                ground_truth_control=optimal_control)
    simulate._SEED_COUNTER[0] = c0
    ws = simulate.rollout(sde, xb.repeat(6, 1), solver.ts, st.lmbd)
    lw = (ws.lw[0] + ws.lw[1] + ws.lw[2])
    wts = torch.exp(lw)
    assert abs(float(nc) - float(wts.mean())) <= 1e-6 * float(wts.mean())
    assert abs(float(nc_err) - float(wts.std() / math.sqrt(wts.numel() - 1))) <= 1e-5 * float(nc_err)
    gt = optimal_control(solver.ts, ws.states)[:-1]
    want_sqd = torch.sum((gt - ws.controls) ** 2 * wts.reshape(1, -1, 1)) / (K * B) / 6
    assert abs(float(sqd) - float(want_sqd)) <= 1e-5 * abs(float(want_sqd))
    # weights-only mode draws the same paths: same constant without a ground-truth control
    simulate._SEED_COUNTER[0] = c0
    nc2, _, none = sb.normalization_constant(sde, xb, solver.ts, cfg, n_batches_normalization=6)
    assert none is None and abs(float(nc2) - float(nc)) <= 1e-6 * float(nc)
