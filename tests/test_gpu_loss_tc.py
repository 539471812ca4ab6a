"""tcgen05 K3 (csrc/loss_tc.cu + csrc/wgrad_tc.cu), called through the C ABI with SOCM_LOSS_FORCE_TC.

The forward pass is 3xTF32 (4e-7 from fp32 per network evaluation), so a pre-activation that is within ~1e-6 of zero
can get the other ReLU mask than in an fp32 evaluation -- and that unit's whole gradient contribution flips (one flip
moves a gradient tensor of a few-thousand-point batch by ~1e-3; the same happens between any two fp32 evaluations,
only 4x less often).  Every comparison below therefore runs on KINK-FREE inputs: points / paths that keep a safe
distance from every kink (oracle/socm_oracle.py:kink_free_attempts, tests/helpers.py:gpu_kink_free_noise), where the
gradient is well defined and the north star's tolerances -- loss 1e-5..1e-4, every gradient tensor 1e-4 -- hold."""
import pytest
import torch
import torch.nn.functional as F

from helpers import gpu_kink_free_noise, make_product_sde, orc, random_setting, rel_l2, seeded_mnet, seeded_unet

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(params=["f16", "tf32"], autouse=True)
def engine(request):
    """Every test of this file runs on both tensor-core engines: the fp16-split one with two CTAs per SM
    (csrc/unet_h.cuh, the default for d <= 15) and the 3xTF32 one (csrc/unet_tc.cuh)."""
    from soc_matching_b200 import simulate
    simulate.ENGINE = request.param
    yield request.param
    simulate.ENGINE = None
    from soc_matching_b200 import _lib
    simulate.sync_engine(_lib.load())   # the library-side default follows (calls without an engine flag)


NAMES = ["down_0", "down_1", "down_2", "res_0", "res_1", "res_2", "up_2", "up_1", "up_0"]


def unet64(P, tx):
    lin = lambda n, v: F.linear(v, P[n + ".0.weight"], P[n + ".0.bias"])  # noqa: E731
    z1 = lin("down_0", tx); r1 = torch.relu(z1)
    z2 = lin("down_1", r1); r2 = torch.relu(z2)
    z3 = lin("down_2", r2); r3 = torch.relu(z3)
    y2 = lin("up_2", r3); o2 = torch.relu(y2) + lin("res_2", r2)
    y1 = lin("up_1", o2); o1 = torch.relu(y1) + lin("res_1", r1)
    y0 = lin("up_0", o1)
    return torch.relu(y0) + lin("res_0", tx), (z1, z2, z3, y2, y1, y0)


@pytest.mark.parametrize("d,K,B", [(10, 40, 300), (1, 30, 129), (20, 9, 200)])
def test_k3_tc_matches_fp64_autograd_away_from_kinks(d, K, B, engine):
    _k3_vs_fp64(d, K, B, engine)


@pytest.mark.parametrize("x_scale,t_scale,w_sigma", [(30.0, 1.0, 0.3), (0.03, 1.0, 0.3), (1.0, 1e3, 0.3), (1.0, 1e-4, 0.3),
                                                      (1.0, 1.0, 3.0), (1.0, 1.0, None)])
def test_k3_tc_dynamic_range(x_scale, t_scale, w_sigma, engine):
    """The fp16-split engine scales every operand by calibrated powers of two: states far from / close to the origin,
    targets (hence loss gradients) three orders of magnitude up and four down, importance weights spread over e^{+-9},
    and all-zero weights (w_sigma None: loss and every gradient exactly zero, nothing non-finite)."""
    _k3_vs_fp64(10, 20, 200, engine, x_scale, t_scale, w_sigma)


def _k3_vs_fp64(d, K, B, engine, x_scale=1.0, t_scale=1.0, w_sigma=0.3):
    from soc_matching_b200 import _lib, networks
    lib = _lib.load()
    p = {k: v.to(DEV) for k, v in seeded_unet(d, [256, 128, 64], 31 + d).items()}
    unet = networks.FullyConnectedUNet(d, (256, 128, 64), 1.0).to(DEV)
    unet.load_state_dict(p)
    udesc, keep = networks.unet_desc(unet)
    g = torch.Generator(DEV).manual_seed(d)
    ts = torch.linspace(0, 1, K + 1, device=DEV)
    P = {k: v.double().requires_grad_(True) for k, v in p.items()}
    states = x_scale * torch.randn(K + 1, B, d, device=DEV, generator=g)
    for _ in range(20):   # move every point that sits within 1e-4 of a ReLU kink
        tx = torch.cat([ts.reshape(-1, 1, 1).expand(K + 1, B, 1), states], -1).double()
        with torch.no_grad():
            _, pre = unet64(P, tx)
        near = torch.stack([(z.abs() < 1e-4).any(-1) for z in pre]).any(0)
        if not bool(near.any()):
            break
        states[near] = x_scale * torch.randn(int(near.sum()), d, device=DEV, generator=g)
    assert not bool(near.any())
    ldt = ((K + 1) * d + 3) // 4 * 4
    target = t_scale * torch.randn(B, ldt, device=DEV, generator=g)
    w = (torch.exp(w_sigma * torch.randn(B, device=DEV, generator=g)) if w_sigma is not None
         else torch.zeros(B, device=DEV))
    G = torch.zeros(B, ldt, device=DEV)
    grad = torch.zeros(int(lib.socm_unet_param_count(udesc)), device=DEV)
    loss = torch.zeros(1, device=DEV, dtype=torch.float64)
    ws = torch.zeros(int(lib.socm_loss_workspace_bytes(udesc, B, K)) // 4 + 1024, device=DEV)
    st = _lib.Setting()
    eye, kap = torch.eye(d, device=DEV), torch.ones(d, device=DEV)
    st.kind, st.d, st.sigma_is_identity, st.lmbd = 2, d, 1, 1.0
    st.sigma, st.sigma_inv, st.kappa, st.nu = eye.data_ptr(), eye.data_ptr(), kap.data_ptr(), kap.data_ptr()
    scale = 1.0 / ((K + 1) * B)
    _lib.check(lib.socm_unet_loss_fwdbwd_f32(st, udesc, None, ts.data_ptr(), states.data_ptr(), target.data_ptr(),
                                             ldt, w.data_ptr(), None, scale, B, K, G.data_ptr(), grad.data_ptr(),
                                             loss.data_ptr(), ws.data_ptr(),
                                             _lib.LOSS_FORCE_TC | (_lib.LOSS_F16 if engine == "f16" else _lib.LOSS_TF32),
                                             _lib.stream_ptr()))
    torch.cuda.synchronize()
    if w_sigma is None:
        assert float(loss) == 0.0 and float(grad.abs().max()) == 0.0 and float(G.abs().max()) == 0.0
        return
    tx = torch.cat([ts.reshape(-1, 1, 1).expand(K + 1, B, 1), states], -1).double()
    out, _ = unet64(P, tx)
    out.retain_grad()
    tgt = target[:, :(K + 1) * d].reshape(B, K + 1, d).permute(1, 0, 2).double()
    L = (((out - tgt) ** 2).sum(-1) * w.double()[None]).sum() * scale
    L.backward()
    assert abs(float(loss) - float(L)) <= 2e-5 * abs(float(L))
    Gwant = -out.grad.permute(1, 0, 2).reshape(B, (K + 1) * d)           # G = d loss / d target
    assert rel_l2(G[:, :(K + 1) * d].cpu(), Gwant.cpu()) <= 2e-5
    o, gc = 0, grad.cpu()
    for n in NAMES:
        for suffix in (".0.weight", ".0.bias"):
            t = P[n + suffix].grad.cpu()
            got = gc[o:o + t.numel()].reshape(t.shape)
            o += t.numel()
            assert rel_l2(got, t) <= 1e-4, (n + suffix, rel_l2(got, t))
            assert rel_l2(got, t) <= 2e-5, (n + suffix, rel_l2(got, t))   # measured: 2e-6 .. 6e-6
    del keep


@pytest.mark.parametrize("kind,d,K,B", [("double_well", 10, 60, 128), ("ou_quadratic", 20, 12, 40)])
def test_socm_iteration_with_tc_k3_matches_oracle(kind, d, K, B):
    """Public API end to end (tcgen05 rollout + tcgen05 K3) against the oracle, loss / gradients 1e-4."""
    import soc_matching_b200 as sb
    st = random_setting(kind, d, seed=d + K)
    hd, hm = [256, 128, 64], [128, 128]
    unet, mnet = seeded_unet(d, hd, 21 + d), seeded_mnet(d, hm, 22 + d, 0.1)
    gam = {"gamma": torch.tensor([2.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    x0 = torch.zeros(d) if kind == "double_well" else 0.3 * torch.ones(d)
    ts = torch.linspace(0, 1.0, K + 1)
    noises = torch.randn(K, B, d, generator=torch.Generator().manual_seed(5))
    torch.set_num_threads(8)
    traj = orc.rollout(st, unet, x0.repeat(B, 1), ts, noises=noises)
    pu = {k: v.clone().requires_grad_(True) for k, v in unet.items()}
    pm = {k: v.clone().requires_grad_(True) for k, v in mnet.items()}
    pg = {k: v.clone().requires_grad_(True) for k, v in gam.items()}
    obj, wm, _ = orc.socm_loss(st, pu, pm, pg, ts, traj, algorithm="SOCM")
    obj.backward()
    sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
    solver = sb.SOC_Solver(sde, x0.to(DEV), None, T=1.0, num_steps=K, lmbd=st.lmbd, d=d, sigma=sde.sigma)
    solver.force_tc = True
    solver.inject_noise(noises.to(DEV))
    out = solver.loss(B, algorithm="SOCM")
    out[0].backward()
    assert abs(float(out[0].detach()) - float(obj.detach())) <= 1e-4 * abs(float(obj.detach()))
    for n, prm in sde.nabla_V.named_parameters():
        assert rel_l2(prm.grad.cpu(), pu[n].grad) <= 1e-4, (n, rel_l2(prm.grad.cpu(), pu[n].grad))
    for n, prm in sde.M.sigmoid_layers.named_parameters():
        assert rel_l2(prm.grad.cpu(), pm["sigmoid_layers." + n].grad) <= 1e-4, n
    assert rel_l2(sde.gamma.grad.cpu(), pg["gamma"].grad) <= 1e-4


def test_tc_k3_is_the_default_for_large_batches_and_agrees_with_ffma():
    """(K+1)*B >= SOCM_LOSS_TC_MIN_POINTS dispatches to the tcgen05 kernels: loss 1e-5 and every gradient tensor
    1e-4 against the fp32 FFMA kernel on the same kink-free rollout."""
    import soc_matching_b200 as sb
    d, K, B = 10, 63, 1024
    st = random_setting("double_well", d, seed=4)
    hd, hm = [256, 128, 64], [128, 128]
    unet, mnet = seeded_unet(d, hd, 5), seeded_mnet(d, hm, 6, 0.1)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
    noises = gpu_kink_free_noise(sde, torch.zeros(d, device=DEV), torch.linspace(0, 1.0, K + 1, device=DEV), B, seed=2)
    res = []
    for ffma in (True, False):
        sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
        solver = sb.SOC_Solver(sde, torch.zeros(d, device=DEV), None, T=1.0, num_steps=K, lmbd=1.0, d=d,
                               sigma=sde.sigma)
        solver.force_ffma = ffma
        solver.inject_noise(noises)
        out = solver.loss(B, algorithm="SOCM")
        out[0].backward()
        res.append((float(out[0].detach()), {n: q.grad.clone() for n, q in sde.named_parameters() if q.grad is not None}))
    (l0, g0), (l1, g1) = res
    assert abs(l0 - l1) <= 1e-5 * abs(l0)
    for n in g0:
        assert rel_l2(g1[n], g0[n]) <= 1e-4, (n, rel_l2(g1[n], g0[n]))


@pytest.mark.parametrize("d,K,B", [(10, 200, 300), (1, 150, 64), (20, 50, 129), (3, 7, 5)])
def test_target_gemm_tc_matches_fp64(d, K, B, engine):
    """K2 forward on tcgen05 (csrc/target_h.cu: fp16 hi/lo split on kind::f16; csrc/target_tc.cu: 3xTF32): target = R L^T
    with the block-triangular L, against torch fp64 (segmented accumulation: 6e-6 relative; the fp32 SIMT kernel: 2e-6).
    The operands span orders of magnitude (per-path scales on R, per-column scales on L)."""
    from soc_matching_b200 import _lib, simulate
    lib = _lib.load()
    simulate.sync_engine(lib)   # the `engine` fixture set simulate.ENGINE: this entry point has no engine flag
    g = torch.Generator(DEV).manual_seed(K)
    nrows, kdim = (K + 1) * d, (2 * K + 1) * d
    ldr, ldt = (kdim + 3) // 4 * 4, (nrows + 3) // 4 * 4
    L = torch.randn(nrows, ldr, device=DEV, generator=g)
    i_of_row = torch.arange(nrows, device=DEV) // d
    col = torch.arange(ldr, device=DEV)
    L[(col[None, :] < 2 * i_of_row[:, None] * d) | (col[None, :] >= kdim)] = 0.0     # structure of mtable.build_L
    L *= torch.exp(torch.randn(1, ldr, device=DEV, generator=g)) * 1e-2
    R = torch.randn(B, ldr, device=DEV, generator=g) * torch.exp(2.0 * torch.randn(B, 1, device=DEV, generator=g)) * 50.0
    R[:, kdim:] = 0.0
    T = torch.full((B, ldt), float("nan"), device=DEV)
    ws = torch.empty(int(lib.socm_target_gemm_tc_workspace_bytes(K, d)), device=DEV, dtype=torch.uint8)
    _lib.check(lib.socm_target_gemm_tc_f32(L.data_ptr(), R.data_ptr(), B, K, d, ldr, T.data_ptr(), ldt, ws.data_ptr(),
                                           _lib.stream_ptr()))
    T2 = torch.empty(B, ldt, device=DEV)
    _lib.check(lib.socm_target_gemm_f32(L.data_ptr(), R.data_ptr(), B, K, d, ldr, T2.data_ptr(), ldt, _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = R.double() @ L.double().t()
    assert rel_l2(T[:, :nrows].cpu(), want.cpu()) <= 6e-6      # measured 3.5e-6 at K = 200 (segmented accumulation)
    assert rel_l2(T2[:, :nrows].cpu(), want.cpu()) <= 2e-6


@pytest.mark.parametrize("d,K,B", [(10, 200, 300), (1, 150, 64), (20, 50, 129), (3, 7, 5), (10, 100, 2048)])
def test_target_gemm_bwd_tc_matches_fp64(d, K, B, engine):
    """K2 backward on tcgen05 (csrc/target_bwd_h.cu: fp16 hi/lo planes on kind::f16; csrc/target_bwd_tc.cu: 3xTF32):
    dL += G^T R on the block-upper-triangular part, against torch fp64 (segmented accumulation: 1e-5 relative) and the
    fp32 SIMT kernel.  The operands span several orders of magnitude (per-path weights on G, per-column scales on R):
    the fp16 engine scales by the exact maxima."""
    from soc_matching_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(DEV).manual_seed(K + B)
    nrows, kdim = (K + 1) * d, (2 * K + 1) * d
    ldr, ldt = (kdim + 3) // 4 * 4, (nrows + 3) // 4 * 4
    G = torch.randn(B, ldt, device=DEV, generator=g) * torch.exp(2.0 * torch.randn(B, 1, device=DEV, generator=g)) * 1e-3
    R = torch.randn(B, ldr, device=DEV, generator=g) * torch.exp(torch.randn(1, ldr, device=DEV, generator=g)) * 30.0
    base = torch.randn(nrows, ldr, device=DEV, generator=g)
    i_of_row = torch.arange(nrows, device=DEV) // d
    col = torch.arange(ldr, device=DEV)
    keep = (col[None, :] >= 2 * i_of_row[:, None] * d) & (col[None, :] < kdim)
    dL = base.clone()
    ws = torch.empty(int(lib.socm_target_gemm_bwd_tc_workspace_bytes(B, K, d)), device=DEV, dtype=torch.uint8)
    eng = _lib.TARGET_BWD_F16 if engine == "f16" else _lib.TARGET_BWD_TF32
    _lib.check(lib.socm_target_gemm_bwd_tc_f32(G.data_ptr(), R.data_ptr(), B, K, d, ldr, ldt, dL.data_ptr(), 1 | eng,
                                               ws.data_ptr(), _lib.stream_ptr()))
    dL2 = base.clone()
    _lib.check(lib.socm_target_gemm_bwd_f32(G.data_ptr(), R.data_ptr(), B, K, d, ldr, ldt, dL2.data_ptr(), 1,
                                            _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = G[:, :nrows].double().t() @ R.double()
    want = torch.where(keep, want, torch.zeros_like(want))
    assert torch.equal((dL - base)[~keep], torch.zeros_like(base)[~keep])          # structural zeros untouched
    assert rel_l2((dL - base).cpu(), want.cpu()) <= 1e-5
    assert rel_l2((dL2 - base).cpu(), want.cpu()) <= 1e-5
    # without the accumulate bit the result replaces dL
    dL3 = base.clone()
    _lib.check(lib.socm_target_gemm_bwd_tc_f32(G.data_ptr(), R.data_ptr(), B, K, d, ldr, ldt, dL3.data_ptr(), eng,
                                               ws.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert rel_l2(dL3.cpu(), want.cpu()) <= 1e-5


@pytest.mark.parametrize("kind,d,K,B,dense,stopping", [("ou_linear", 10, 40, 257, True, False),
                                                        ("ou_quadratic", 20, 20, 200, True, False),
                                                        ("molecular_dynamics", 1, 150, 200, False, True)])
def test_tc_k3_general_loss_paths_agree_with_ffma(kind, d, K, B, dense, stopping):
    """Dense sigma (generic per-point loss inside K3a) and stopping-time masks (per-point stop indicator,
    Z = sum of indicators) through the tcgen05 K3: loss 1e-5 and every gradient tensor 1e-4 against the fp32 FFMA
    kernels on the same kink-free injected noise."""
    import soc_matching_b200 as sb
    st = random_setting(kind, d, seed=d + K, dense_sigma=dense)
    hd, hm = [256, 128, 64], [64, 64]
    unet = seeded_unet(d, hd, 41 + d, 0.5 if stopping else 1.0)
    mnet = seeded_mnet(d, hm, 42 + d, 0.1, 3 if stopping else 2)
    gam = {"gamma": torch.tensor([2.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    x0 = -torch.ones(d) if stopping else 0.3 * torch.ones(d)
    sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV, stopping=stopping)
    noises = gpu_kink_free_noise(sde, x0.to(DEV), torch.linspace(0, 1.0, K + 1, device=DEV), B, seed=7)
    res = []
    for tc_path in (False, True):
        sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV, stopping=stopping)
        solver = sb.SOC_Solver(sde, x0.to(DEV), None, T=1.0, num_steps=K, lmbd=st.lmbd, d=d, sigma=sde.sigma)
        solver.force_ffma, solver.force_tc = (not tc_path), tc_path
        solver.inject_noise(noises)
        out = solver.loss(B, algorithm="SOCM", use_stopping_time=stopping)
        out[0].backward()
        res.append((float(out[0].detach()), {n: q.grad.clone() for n, q in sde.nabla_V.named_parameters()},
                    out[7].clone()))
    (l0, g0, s0), (l1, g1, s1) = res
    assert torch.equal(s0, s1)                                   # stopping indicators (tcgen05 vs FFMA rollout)
    assert abs(l0 - l1) <= 1e-5 * abs(l0), (l0, l1)
    for n in g0:
        assert rel_l2(g1[n], g0[n]) <= 1e-4, (n, rel_l2(g1[n], g0[n]))


def test_full_chunk_tc_iteration_agrees_with_fp32_ffma():
    """BASELINE config 5 at the size of one bench chunk (double_well d=10, K=200, B=75 776 paths = 15.2 M
    trajectory points): the tcgen05 path (rollout, target GEMMs, K3) against the exact-fp32 FFMA / SIMT path on the
    same kink-free injected noise.  Loss 1e-5, every gradient tensor 1e-4 (north star)."""
    import soc_matching_b200 as sb
    from soc_matching_b200 import simulate
    d, K, B = 10, 200, 75776
    st = random_setting("double_well", d, seed=4)
    hd, hm = [256, 128, 64], [128, 128]
    unet, mnet = seeded_unet(d, hd, 5), seeded_mnet(d, hm, 6, 0.1)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    res = []
    sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
    noises = gpu_kink_free_noise(sde, torch.zeros(d, device=DEV), torch.linspace(0, 1.0, K + 1, device=DEV), B, seed=77)
    del sde
    for ffma in (True, False):
        sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
        solver = sb.SOC_Solver(sde, torch.zeros(d, device=DEV), None, T=1.0, num_steps=K, lmbd=1.0, d=d,
                               sigma=sde.sigma)
        solver.force_ffma = ffma
        solver.inject_noise(noises)
        out = solver.loss(B, algorithm="SOCM")
        out[0].backward()
        res.append((float(out[0].detach()), float(out[5]), out[7].clone(),
                    {n: q.grad.clone() for n, q in sde.named_parameters() if q.grad is not None}))
        del solver, sde, out
        torch.cuda.empty_cache()
    (l0, w0, s0, g0), (l1, w1, s1, g1) = res
    assert abs(l1 - l0) <= 1e-5 * abs(l0), (l0, l1)
    assert abs(w1 - w0) <= 1e-5 * abs(w0)
    assert torch.equal(s0, s1)
    errs = {n: rel_l2(g1[n], g0[n]) for n in g0}
    print("full chunk, tcgen05 vs fp32 FFMA, per-tensor gradient error:", {k: f"{v:.1e}" for k, v in errs.items()})
    for n in g0:
        assert errs[n] <= 1e-4, (n, errs[n])


def test_full_chunk_k3_tc_against_fp64_autograd(engine):
    """K3 on the tensor cores against torch fp64 autograd over all 15.2 M points of a real bench chunk (states of a
    tcgen05 rollout, double_well d=10, K=200, B=75 776; a well-conditioned random target): every gradient tensor
    within 1e-4 (measured 2e-5..4.5e-5; the fp32 FFMA kernel: 1e-7..1.3e-5), loss 1e-6.  This is the error of the
    3xTF32 chain (truncating accumulation, 96 MMAs per 256-wide contraction) plus ReLU-mask flips of
    pre-activations within ~2e-6 of zero; the weight-gradient accumulators are flushed every 8 tiles."""
    from soc_matching_b200 import _lib, networks, simulate
    lib = _lib.load()
    d, K, B = 10, 200, 75776
    st_ = random_setting("double_well", d, seed=4)
    hd, hm = [256, 128, 64], [128, 128]
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(st_, seeded_unet(d, hd, 5), seeded_mnet(d, hm, 6, 0.1), gam, hd, hm, DEV)
    ts = torch.linspace(0, 1, K + 1, device=DEV)
    wsp = simulate.rollout(sde, torch.zeros(B, d, device=DEV), ts, 1.0, seed=99)
    states, unet = wsp.states, sde.nabla_V
    udesc, keep = networks.unet_desc(unet)
    g = torch.Generator(DEV).manual_seed(1)
    ldt = ((K + 1) * d + 3) // 4 * 4
    target = 3.0 * torch.randn(B, ldt, device=DEV, generator=g)
    w = torch.exp(wsp.lw[0] + wsp.lw[1] + wsp.lw[2])
    st = _lib.Setting()
    eye, kap = torch.eye(d, device=DEV), torch.ones(d, device=DEV)
    st.kind, st.d, st.sigma_is_identity, st.lmbd = 2, d, 1, 1.0
    st.sigma, st.sigma_inv, st.kappa, st.nu = eye.data_ptr(), eye.data_ptr(), kap.data_ptr(), kap.data_ptr()
    ws = torch.zeros(int(lib.socm_loss_workspace_bytes(udesc, B, K)) // 4 + 1024, device=DEV)
    scale = 1.0 / ((K + 1) * B)
    G = torch.zeros(B, ldt, device=DEV)
    grad = torch.zeros(int(lib.socm_unet_param_count(udesc)), device=DEV)
    loss = torch.zeros(1, device=DEV, dtype=torch.float64)
    _lib.check(lib.socm_unet_loss_fwdbwd_f32(st, udesc, None, ts.data_ptr(), states.data_ptr(), target.data_ptr(), ldt,
                                             w.data_ptr(), None, scale, B, K, G.data_ptr(), grad.data_ptr(),
                                             loss.data_ptr(), ws.data_ptr(),
                                             _lib.LOSS_FORCE_TC | (_lib.LOSS_F16 if engine == "f16" else _lib.LOSS_TF32),
                                             _lib.stream_ptr()))
    torch.cuda.synchronize()
    P = {n + sfx: getattr(getattr(unet, n)[0], sfx[1:]).detach().double().requires_grad_(True)
         for n in NAMES for sfx in (".weight", ".bias")}
    lin = lambda n, v: F.linear(v, P[n + ".weight"], P[n + ".bias"])  # noqa: E731
    tot = 0.0
    for i0 in range(0, K + 1, 8):     # fp64 autograd in slabs of 8 grid times
        i1 = min(K + 1, i0 + 8)
        tx = torch.cat([ts[i0:i1].double().reshape(-1, 1, 1).expand(i1 - i0, B, 1), states[i0:i1].double()], -1)
        r1 = torch.relu(lin("down_0", tx)); r2 = torch.relu(lin("down_1", r1)); r3 = torch.relu(lin("down_2", r2))
        o2 = torch.relu(lin("up_2", r3)) + lin("res_2", r2)
        o1 = torch.relu(lin("up_1", o2)) + lin("res_1", r1)
        outv = torch.relu(lin("up_0", o1)) + lin("res_0", tx)
        tgt = target[:, i0 * d:i1 * d].reshape(B, i1 - i0, d).permute(1, 0, 2).double()
        L = (((outv - tgt) ** 2).sum(-1) * w.double()[None]).sum() * scale
        L.backward()
        tot += float(L.detach())
    assert abs(float(loss) - tot) <= 1e-6 * abs(tot)
    off = 0
    for n in NAMES:
        for sfx in (".weight", ".bias"):
            t = P[n + sfx].grad.flatten()
            got = grad[off:off + t.numel()].double()
            off += t.numel()
            assert rel_l2(got, t) <= 1e-4, (n + sfx, rel_l2(got, t))
    del keep
