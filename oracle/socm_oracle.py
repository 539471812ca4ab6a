"""CPU oracle for the SOC-matching hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU fp32 restatement of the reference algorithm for the
three-part hot path (Euler-Maruyama rollout -> SOCM matching target -> importance
weighted loss).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product (``soc_matching_b200``) never imports anything from ``oracle/`` and has no
CPU fallback.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference
from ``/root/reference`` (three import stubs), runs it with
seeded inputs and stores its outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.
The reference ships no tests or golden vectors of its own.

Citations are ``file:line`` relative to ``/root/reference``.
All functions work on plain tensors / dicts of tensors; parameter dict keys are the
reference's ``state_dict`` names (e.g. ``down_0.0.weight``) so that weights can be
moved between the reference, this oracle and the product without renaming.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
KINDS = ("ou_quadratic", "ou_linear", "double_well", "molecular_dynamics")


# --------------------------------------------------------------------------------------
# Setting primitives  (SOC_matching/experiment_settings/*.py)
# --------------------------------------------------------------------------------------
@dataclass
class Setting:
    """Closed-form problem data of one benchmark (drift b, costs f/g, sigma, lmbd)."""

    kind: str
    d: int
    sigma: Tensor
    lmbd: float = 1.0
    A: Optional[Tensor] = None      # ou_quadratic, ou_linear
    P: Optional[Tensor] = None      # ou_quadratic
    Q: Optional[Tensor] = None      # ou_quadratic
    omega: Optional[Tensor] = None  # ou_linear
    kappa: Optional[Tensor] = None  # double_well, molecular_dynamics
    nu: Optional[Tensor] = None     # double_well
    T: float = 1.0

    @property
    def has_stopping(self) -> bool:
        # utils.py:33 -- stopping is switched on by the mere presence of ``Phi``.
        return self.kind == "molecular_dynamics"


def _matvec(mat: Tensor, x: Tensor) -> Tensor:
    # The reference writes every d x d product as einsum("ij,...j->...i").
    return torch.einsum("ij,...j->...i", mat, x)


def drift(st: Setting, x: Tensor) -> Tensor:
    """b(x).  OU_quadratic.py:51-52, OU_linear.py:43-44, double_well.py:43-48,
    molecular_dynamics.py:49-53."""
    if st.kind in ("ou_quadratic", "ou_linear"):
        return _matvec(st.A, x)
    kap = st.kappa.reshape((1,) * (x.dim() - 1) + (-1,))
    return -2 * kap * (x**2 - 1) * 2 * x


def grad_drift(st: Setting, x: Tensor) -> Tensor:
    """nabla_b(x), shape x.shape + (d,).  OU_quadratic.py:55-63 (returns A^T),
    double_well.py:51-61 / molecular_dynamics.py:56-66 (diagonal)."""
    if st.kind in ("ou_quadratic", "ou_linear"):
        lead = x.shape[:-1]
        return st.A.t().reshape((1,) * len(lead) + st.A.shape).expand(*lead, st.d, st.d)
    kap = st.kappa.reshape((1,) * (x.dim() - 1) + (-1,))
    return -torch.diag_embed(8 * kap * x**2 + 4 * kap * (x**2 - 1))


def run_cost(st: Setting, x: Tensor) -> Tensor:
    """f(x).  OU_quadratic.py:66-69, OU_linear.py:67-75 (0), double_well.py:64-68 (0),
    molecular_dynamics.py:79-84 (1)."""
    if st.kind == "ou_quadratic":
        return torch.sum(x * _matvec(st.P, x), -1)
    if st.kind == "molecular_dynamics":
        return torch.ones_like(x[..., 0])
    return torch.zeros_like(x[..., 0])


def grad_run_cost(st: Setting, x: Tensor) -> Tensor:
    """nabla_f(x).  OU_quadratic.py:72-73; zero elsewhere."""
    if st.kind == "ou_quadratic":
        return 2 * _matvec(st.P, x)
    return torch.zeros_like(x)


def term_cost(st: Setting, x: Tensor) -> Tensor:
    """g(x).  OU_quadratic.py:76-79, OU_linear.py:81-82, double_well.py:75-84,
    molecular_dynamics.py:69-74 (0)."""
    if st.kind == "ou_quadratic":
        return torch.sum(x * _matvec(st.Q, x), -1)
    if st.kind == "ou_linear":
        return torch.einsum("j,...j->...", st.omega, x)
    if st.kind == "double_well":
        nu = st.nu.reshape((1,) * (x.dim() - 1) + (-1,))
        return torch.sum(nu * (x**2 - 1) ** 2, dim=-1)
    return torch.zeros_like(x[..., 0])


def grad_term_cost(st: Setting, x: Tensor) -> Tensor:
    """nabla_g(x).  OU_quadratic.py:82-83, OU_linear.py:85-96, double_well.py:87-97,
    molecular_dynamics.py:76-77."""
    if st.kind == "ou_quadratic":
        return 2 * _matvec(st.Q, x)
    if st.kind == "ou_linear":
        return st.omega.reshape((1,) * (x.dim() - 1) + (-1,)).expand_as(x).clone()
    if st.kind == "double_well":
        nu = st.nu.reshape((1,) * (x.dim() - 1) + (-1,))
        return 2 * nu * (x**2 - 1) * 2 * x
    return torch.zeros_like(x)


def stop_fn(st: Setting, x: Tensor) -> Tensor:
    """Phi(x) = -x_0; the process stops once Phi <= 0.  molecular_dynamics.py:91-95."""
    return -x[..., 0]


# --------------------------------------------------------------------------------------
# Networks  (SOC_matching/models.py)
# --------------------------------------------------------------------------------------
def unet_apply(p: Dict[str, Tensor], tx: Tensor) -> Tensor:
    """FullyConnectedUNet.forward, models.py:233-242 (ReLU also on the last up layer,
    models.py:228)."""

    def lin(name: str, v: Tensor) -> Tensor:
        return F.linear(v, p[name + ".0.weight"], p[name + ".0.bias"])

    r1 = torch.relu(lin("down_0", tx))
    r2 = torch.relu(lin("down_1", r1))
    r3 = torch.relu(lin("down_2", r2))
    o2 = torch.relu(lin("up_2", r3)) + lin("res_2", r2)
    o1 = torch.relu(lin("up_1", o2)) + lin("res_1", r1)
    return torch.relu(lin("up_0", o1)) + lin("res_0", tx)


def _mlp3(p: Dict[str, Tensor], v: Tensor) -> Tensor:
    h = torch.relu(F.linear(v, p["sigmoid_layers.0.weight"], p["sigmoid_layers.0.bias"]))
    h = torch.relu(F.linear(h, p["sigmoid_layers.2.weight"], p["sigmoid_layers.2.bias"]))
    return F.linear(h, p["sigmoid_layers.4.weight"], p["sigmoid_layers.4.bias"])


def m_apply(p: Dict[str, Tensor], gamma: Tensor, t: Tensor, s: Tensor, d: int) -> Tensor:
    """SigmoidMLP.forward, models.py:265-275:  M = e^{-g(s-t)} I + (1-e^{-g(s-t)}) N(t,s)."""
    ts = torch.cat((t.unsqueeze(1), s.unsqueeze(1)), dim=1)
    net = _mlp3(p, ts).reshape(-1, d, d)
    grow = torch.exp(gamma * (ts[:, 1] - ts[:, 0])).unsqueeze(1).unsqueeze(2)
    eye = torch.eye(d, dtype=ts.dtype).unsqueeze(0)
    return (1 / grow) * eye.repeat(ts.shape[0], 1, 1) + (1 - 1 / grow) * net


def m_apply_stopping(
    p: Dict[str, Tensor], gamma: Tensor, gamma2: Tensor, gamma3: Tensor,
    t: Tensor, s: Tensor, tau: Tensor, d: int, T: float = 1.0,
) -> Tensor:
    """TwoBoundarySigmoidMLP.forward, models.py:311-393.  t, s: (P,), tau: (P, B)
    -> (P, B, d, d)."""
    zeros = torch.zeros_like(s).unsqueeze(1)
    ones = torch.ones_like(s).unsqueeze(1)
    net_stopped = _mlp3(p, torch.cat((t.unsqueeze(1), s.unsqueeze(1), zeros), 1)).reshape(-1, 1, d, d)
    net_alive = _mlp3(p, torch.cat((t.unsqueeze(1), s.unsqueeze(1), ones), 1)).reshape(-1, 1, d, d)
    eye = torch.eye(d).unsqueeze(0).unsqueeze(0)

    ratio = (1 - torch.exp(-gamma * (s - t))).unsqueeze(1) / (
        1 - torch.exp(-gamma * torch.abs(tau - t.unsqueeze(1))) + 1e-7
    )
    fac1 = torch.nan_to_num(1 - torch.minimum(ratio, torch.tensor([1])), nan=0.0)
    fac1 = fac1 * (tau - 1e-3 > s.unsqueeze(1)).to(torch.int)

    decay3 = torch.exp(-gamma3 * (s - t)).unsqueeze(1)
    alive = (tau > T - 1e-3).to(torch.int)
    part_id = ((1 - alive) * fac1 + alive * decay3).unsqueeze(2).unsqueeze(3) * eye.repeat(
        t.shape[0], alive.shape[1], 1, 1
    )

    def bump(v: Tensor) -> Tensor:
        return (1 - torch.exp(-gamma2 * v)) * (torch.exp(-gamma2 * v) - torch.exp(-gamma2))

    part_net = ((1 - alive) * bump(fac1)).unsqueeze(2).unsqueeze(3) * net_stopped + (
        alive * (1 - decay3)
    ).unsqueeze(2).unsqueeze(3) * net_alive
    return part_id + part_net


# --------------------------------------------------------------------------------------
# Warm-start control as a per-grid-time affine table
# --------------------------------------------------------------------------------------
@dataclass
class WarmStartTable:
    """u_ws(t_k, x) = sigma^{-1} ( c_k + A_k x - b(x) ).

    models.py:163-199 evaluates the Gaussian-path spline drift
    ``dmean + a (x - mean)`` (gsbm_lib.py:227-306) minus the base drift, times
    sigma^{-1}.  With the spline frozen, the drift is affine in x with coefficients
    that depend on the (shifted) grid time only, so it is tabulated once:
    ``A_roll/c_roll`` at the rank-2 branch's shifted times (rollout, K rows) and
    ``A_loss/c_loss`` at the rank-3 branch's shifted times (loss, K+1 rows); the two
    differ at t = T/2."""

    A_roll: Tensor  # (K, d, d)
    c_roll: Tensor  # (K, d)
    A_loss: Tensor  # (K+1, d, d)
    c_loss: Tensor  # (K+1, d)


def warm_start_eval(st: Setting, A: Tensor, c: Tensor, x: Tensor) -> Tensor:
    """x: (..., d) with A: (d, d), c: (d,) broadcast, or x: (K+1, B, d) with A: (K+1, d, d)."""
    if A.dim() == 2:
        aff = c + _matvec(A, x)
    else:
        aff = c.unsqueeze(1) + torch.einsum("kij,kbj->kbi", A, x)
    return _matvec(torch.inverse(st.sigma), aff - drift(st, x))


# --------------------------------------------------------------------------------------
# Rollout  (SOC_matching/utils.py:17-128)
# --------------------------------------------------------------------------------------
def learned_control(st: Setting, unet: Dict[str, Tensor], t0: Tensor, x: Tensor) -> Tensor:
    """NeuralSDE.control rank-2 branch without warm start, method.py:64-72."""
    tcol = t0.reshape(-1, 1).expand(x.shape[0], 1)
    grad_v = unet_apply(unet, torch.cat([tcol, x], dim=-1)).reshape(x.shape)
    return -torch.einsum("ij,bj->bi", st.sigma.t(), grad_v)


def rollout(
    st: Setting,
    unet: Dict[str, Tensor],
    x0: Tensor,
    ts: Tensor,
    noises: Optional[Tensor] = None,
    warm: Optional[WarmStartTable] = None,
    control_fn: Optional[Callable[[int, Tensor, Tensor], Tensor]] = None,
):
    """Batched Euler-Maruyama with control, importance-weight accumulation and the
    stopping-time logic; restates utils.py:17-128 step by step (same op order).

    x0: (B, d); ts: (K+1,); noises: optional injected (K, B, d) (otherwise drawn with
    ``torch.randn_like`` exactly like utils.py:39).  Returns the reference's 8-tuple:
    states (K+1,B,d), noises (K,B,d), stop_indicators (K+1,B), fractional_timesteps
    (K,B), logw_det (B,), logw_sto (B,), logw_term (B,), controls (K,B,d)."""
    lmbd = st.lmbd
    B = x0.shape[0]
    x = x0
    xs, eps_all, us = [x0], [], []
    alive_all = [torch.ones(B)]
    fracs = []
    lw_det = torch.zeros(B)
    lw_sto = torch.zeros(B)
    alive = torch.ones(B)
    for k, (t0, t1) in enumerate(zip(ts[:-1], ts[1:])):
        dt = t1 - t0                                               # utils.py:38
        eps = torch.randn_like(x) if noises is None else noises[k]  # utils.py:39
        eps_all.append(eps)
        if control_fn is not None:
            u = control_fn(k, t0, x)
        else:
            u = learned_control(st, unet, t0, x)                   # utils.py:41
            if warm is not None:                                   # method.py:77-78
                u = u + warm_start_eval(st, warm.A_roll[k], warm.c_roll[k], x)
        if st.has_stopping:
            phi_before = stop_fn(st, x)                            # utils.py:43
            x_before = x
        step = (drift(st, x) + torch.einsum("ij,bj->bi", st.sigma, u)) * dt + torch.sqrt(
            lmbd * dt
        ) * torch.einsum("ij,bj->bi", st.sigma, eps)               # utils.py:45-47
        x = x + alive.unsqueeze(1) * step                          # utils.py:48
        if st.has_stopping:
            phi_after = stop_fn(st, x)                             # utils.py:50
            still = torch.logical_and(phi_before > 0, phi_after > 0).to(torch.float)
            crossed = torch.logical_and(phi_before > 0, phi_after < 0).to(torch.float)
            frac = crossed * (phi_before / (phi_before - phi_after + 1e-6) + 1e-6)  # :57-61
            x = crossed.unsqueeze(1) * (
                x_before + frac.unsqueeze(1) * alive.unsqueeze(1) * step
            ) + (1 - crossed.unsqueeze(1)) * x                     # utils.py:62-69
            eff_dt = crossed * frac**2 * dt + still * dt           # utils.py:70-72
            fracs.append(eff_dt)
            alive = stop_fn(st, x) > 0                             # utils.py:74
            alive_all.append(alive)
        else:
            eff_dt = dt
            fracs.append(dt * torch.ones(B))                       # utils.py:77
            alive_all.append(torch.ones(B))
        xs.append(x)
        us.append(u)
        # running cost at the post-update state (utils.py:48 rebinding, :87/:95)
        lw_det = lw_det + eff_dt / lmbd * (-run_cost(st, x) - 0.5 * torch.sum(u**2, dim=1))
        lw_sto = lw_sto + torch.sqrt(eff_dt / lmbd) * (-torch.sum(u * eps, dim=1))
    lw_term = -term_cost(st, x) / lmbd                             # utils.py:101
    return (
        torch.stack(xs).detach(),
        torch.stack(eps_all).detach(),
        torch.stack(alive_all).detach(),
        torch.stack(fracs).detach(),
        lw_det.detach(),
        lw_sto.detach(),
        lw_term.detach(),
        torch.stack(us).detach(),
    )


# --------------------------------------------------------------------------------------
# SOCM loss  (SOC_matching/method.py:223-287, 289-369, 480-720, 897-906)
# --------------------------------------------------------------------------------------
def _pair_grid(ts: Tensor, T: float):
    """(t, s) pairs with s >= t; s regenerated per row by linspace, method.py:535-547."""
    K = ts.shape[0] - 1
    s_rows, t_rows = [], []
    for k, t in enumerate(ts):
        s_rows.append(torch.linspace(t, T, K + 1 - k))
        t_rows.append(t * torch.ones(K + 1 - k))
    return torch.cat(t_rows), torch.cat(s_rows)


def nabla_v_all(st: Setting, unet: Dict[str, Tensor], ts: Tensor, states: Tensor,
                warm: Optional[WarmStartTable]) -> Tensor:
    """UNet at all (K+1) B points [minus sigma^{-T} u_ws], method.py:272-287."""
    cols = ts.unsqueeze(1).unsqueeze(2).repeat(1, states.shape[1], 1)
    tx = torch.cat([cols, states], dim=-1).reshape(-1, states.shape[2] + 1)
    gv = unet_apply(unet, tx).reshape(states.shape)
    if warm is not None:
        uws = warm_start_eval(st, warm.A_loss, warm.c_loss, states).detach()
        gv = gv - torch.einsum("ij,abj->abi", torch.inverse(st.sigma).t(), uws)
    return gv


def socm_loss(
    st: Setting,
    unet: Dict[str, Tensor],
    mnet: Dict[str, Tensor],
    gammas: Dict[str, Tensor],
    ts: Tensor,
    traj,
    algorithm: str = "SOCM",
    warm: Optional[WarmStartTable] = None,
    use_stopping_time: bool = False,
    add_weights: bool = False,
    y0: Optional[Tensor] = None,
):
    """Faithful restatement of SOC_Solver.loss for algorithm in {SOCM, SOCM_const_M, SOCM_adjoint,
    cross_entropy, variance, log-variance, moment}
    given a finished rollout ``traj`` (the 8-tuple of :func:`rollout`).  The SOCM
    branch keeps the reference's structure on purpose -- reverse-mode ``jacrev`` for
    d/ds M and the materialised (K+1, K+1, B, d, d) integrand -- because this function
    is also what ``bench.py`` times as the CPU baseline.

    Returns (objective, mean(w), std(w)) with autograd to unet / mnet / gammas."""
    states, noises, stop_ind, frac_dt, lw_det, lw_sto, lw_term, controls = traj
    lmbd, d = st.lmbd, st.d
    K1 = ts.shape[0]
    B = states.shape[1]
    weight = torch.exp(lw_det + lw_sto + lw_term)                  # method.py:258-262
    gv = nabla_v_all(st, unet, ts, states, warm)
    sig_inv_t = torch.inverse(st.sigma).t()
    eps_t = torch.einsum("ij,abj->abi", sig_inv_t, noises)
    u_t = torch.einsum("ij,abj->abi", sig_inv_t, controls)

    if algorithm == "SOCM_const_M":                                # method.py:289-369
        gf = grad_run_cost(st, states)[:-1]
        gb = grad_drift(st, states)[:-1]
        term2 = -math.sqrt(lmbd) * torch.einsum("abij,abj->abi", gb, eps_t)
        term3 = -torch.einsum("abij,abj->abi", gb, u_t)
        dts = ts[1:] - ts[:-1]

        def tail_sum(v: Tensor, scale: Tensor) -> Tensor:
            padded = torch.cat((torch.zeros_like(v[0]).unsqueeze(0), v * scale.unsqueeze(1).unsqueeze(2)), 0)
            return torch.sum(padded, dim=0).unsqueeze(0) - torch.cumsum(padded, dim=0)

        target = (
            tail_sum(gf, dts) + tail_sum(term2, torch.sqrt(dts)) + tail_sum(term3, dts)
            + grad_term_cost(st, states[-1]).unsqueeze(0)
        )
        learned = -torch.einsum("ij,...j->...i", st.sigma.t(), gv)
        wanted = -torch.einsum("ij,...j->...i", st.sigma.t(), target)
        obj = torch.sum((learned - wanted) ** 2 * weight.unsqueeze(0).unsqueeze(2)) / (K1 * B)
        return obj, torch.mean(weight), torch.std(weight)

    if algorithm == "SOCM_exp":                                    # method.py:371-478: M_t(s) = exp(-gamma (s - t)) I
        gamma = gammas["gamma"]
        ef = torch.exp(-gamma * ts)
        ident = torch.eye(d)
        gb = grad_drift(st, states)[:-1] + gamma * ident
        term1 = (ef.unsqueeze(1).unsqueeze(2) * grad_run_cost(st, states))[:-1]
        term2 = ef[:-1].unsqueeze(1).unsqueeze(2) * (-math.sqrt(lmbd) * torch.einsum("abij,abj->abi", gb, eps_t))
        term3 = ef[:-1].unsqueeze(1).unsqueeze(2) * (-torch.einsum("abij,abj->abi", gb, u_t))
        terminal = torch.exp(-gamma * (st.T - ts)).unsqueeze(1).unsqueeze(2) * grad_term_cost(st, states[-1]).unsqueeze(0)
        dts = ts[1:] - ts[:-1]

        def tail(v: Tensor, scale: Tensor) -> Tensor:
            padded = torch.cat((torch.zeros_like(v[0]).unsqueeze(0), v * scale.unsqueeze(1).unsqueeze(2)), 0)
            return (1 / ef).unsqueeze(1).unsqueeze(2) * (torch.sum(padded, dim=0).unsqueeze(0) - torch.cumsum(padded, dim=0))

        target = tail(term1, dts) + tail(term2, torch.sqrt(dts)) + tail(term3, dts) + terminal
        learned = -torch.einsum("ij,...j->...i", st.sigma.t(), gv)
        wanted = -torch.einsum("ij,...j->...i", st.sigma.t(), target)
        obj = torch.sum((learned - wanted) ** 2 * weight.unsqueeze(0).unsqueeze(2)) / (K1 * B)
        return obj, torch.mean(weight), torch.std(weight)

    if algorithm in ("cross_entropy", "variance", "log-variance", "moment"):   # method.py:751-856
        # functionals of the per-path sums of a running term that is quadratic in the learned control
        learned = -torch.einsum("ij,abj->abi", st.sigma.t(), gv)
        term1 = -(1 / lmbd) * torch.sum(learned[:-1] * controls, dim=2)
        term2 = (1 / (2 * lmbd)) * torch.sum(learned**2, dim=2)[:-1]
        det = term1 + term2
        if algorithm != "cross_entropy":
            det = det + (-(1 / lmbd) * run_cost(st, states)[:-1])
        sto = -math.sqrt(1 / lmbd) * torch.sum(learned[:-1] * noises, dim=2)
        if use_stopping_time and algorithm != "cross_entropy":
            det = det * stop_ind[:-1]
            sto = sto * stop_ind[:-1]
        if use_stopping_time:
            det_dt, sto_dt = det * frac_dt, sto * torch.sqrt(frac_dt)
        else:
            dts = ts[1:] - ts[:-1]
            det_dt, sto_dt = det * dts.unsqueeze(1), sto * torch.sqrt(dts).unsqueeze(1)
        det_term, sto_term = torch.sum(det_dt, dim=0), torch.sum(sto_dt, dim=0)
        if algorithm == "cross_entropy":
            return torch.mean((det_term + sto_term) * weight), torch.mean(weight), torch.std(weight)
        g_term = -(1 / lmbd) * term_cost(st, states[-1])
        if algorithm == "log-variance":
            sums = det_term + sto_term + g_term
        elif algorithm == "variance":
            sums = torch.exp(det_term + sto_term + g_term)
        else:
            sums = det_term + sto_term + g_term + y0
        w2 = weight if add_weights else torch.ones_like(weight)
        if algorithm == "moment":
            obj = torch.mean(sums**2 * w2)
        else:
            n = sums.shape[0]
            obj = n / (n - 1) * (torch.mean(sums**2 * w2) - torch.mean(sums * w2) ** 2)
        return obj, torch.mean(weight), torch.std(weight)

    if algorithm == "SOCM_adjoint":                                # method.py:722-749
        # backward recursion of the adjoint a_k along every path (trapezoid rule in f' and b'), target = a
        gf = grad_run_cost(st, states)
        gb = grad_drift(st, states)
        dt = st.T / (K1 - 1)                                       # self.dt = T / num_steps (method.py:169)
        a = grad_term_cost(st, states[-1]).clone()
        a_vectors = torch.zeros_like(states)
        a_vectors[-1] = a
        for k in range(1, K1):
            a = a + dt * ((gf[-1 - k] + gf[-k]) / 2
                          + torch.einsum("mkl,ml->mk", (gb[-1 - k] + gb[-k]) / 2, a))
            a_vectors[-1 - k] = a
        learned = -torch.einsum("ij,...j->...i", st.sigma.t(), gv)
        wanted = -torch.einsum("ij,...j->...i", st.sigma.t(), a_vectors)
        obj = torch.sum((learned - wanted) ** 2 * weight.unsqueeze(0).unsqueeze(2)) / (K1 * B)
        return obj, torch.mean(weight), torch.std(weight)

    assert algorithm == "SOCM"
    gamma = gammas["gamma"]
    t_vec, s_vec = _pair_grid(ts, st.T)
    if use_stopping_time:                                          # method.py:484-507, 524-564
        gamma2, gamma3 = gammas["gamma2"], gammas["gamma3"]
        tau = (torch.sum((stop_fn(st, states) > 0).to(torch.int), dim=0) - 1) / (K1 - 1)
        tau_vec = torch.cat([tau.unsqueeze(0).repeat(K1 - k, 1) for k in range(K1)], dim=0)

        def m_of(t, s, tv):
            return m_apply_stopping(mnet, gamma, gamma2, gamma3, t, s, tv, d)

        jac = torch.func.jacrev(lambda t, s, tv: m_of(t, s, tv).sum(dim=0), argnums=1)
        m_all = m_of(t_vec, s_vec, tau_vec)
        dm_all = torch.nan_to_num(jac(t_vec, s_vec, tau_vec).permute(3, 0, 1, 2))
        m_tab = torch.zeros(K1, K1, B, d, d)
        dm_tab = torch.zeros(K1, K1, B, d, d)
    else:                                                          # method.py:509-522, 565-582
        def m_of(t, s):
            return m_apply(mnet, gamma, t, s, d)

        jac = torch.func.jacrev(lambda t, s: m_of(t, s).sum(dim=0), argnums=1)
        m_all = m_of(t_vec, s_vec)
        dm_all = jac(t_vec, s_vec).permute(2, 0, 1)
        m_tab = torch.zeros(K1, K1, d, d)
        dm_tab = torch.zeros(K1, K1, d, d)
    row = 0
    for k in range(K1):
        n = K1 - k
        m_tab[k, k:] = m_all[row:row + n]
        dm_tab[k, k:] = dm_all[row:row + n]
        row += n

    gf = grad_run_cost(st, states)
    gb = grad_drift(st, states)
    if use_stopping_time:                                          # method.py:584-646
        term1 = torch.einsum("ijmkl,jml->ijmk", m_tab, gf)[:, :-1]
        mb = torch.einsum("ijmkl,jmln->ijmkn", m_tab, gb) - dm_tab
        m_last = m_tab[:, -1]
        terminal = torch.einsum("imkl,ml->imk", m_last, grad_term_cost(st, states[-1]))
    else:
        term1 = torch.einsum("ijkl,jml->ijmk", m_tab, gf)[:, :-1]
        mb = torch.einsum("ijkl,jmln->ijmkn", m_tab, gb) - dm_tab.unsqueeze(2)
        m_last = m_tab[:, -1]
        terminal = torch.einsum("ikl,ml->imk", m_last, grad_term_cost(st, states[-1]))
    term2 = -math.sqrt(lmbd) * torch.einsum("ijmkn,jmn->ijmk", mb[:, :-1], eps_t)
    term3 = -torch.einsum("ijmkn,jmn->ijmk", mb[:, :-1], u_t)
    if use_stopping_time:                                          # method.py:648-673
        w1 = frac_dt.unsqueeze(0).unsqueeze(3)
        w2 = torch.sqrt(frac_dt).unsqueeze(0).unsqueeze(3)
    else:
        dts = ts[1:] - ts[:-1]
        w1 = dts.unsqueeze(1).unsqueeze(2).unsqueeze(0)
        w2 = torch.sqrt(dts).unsqueeze(1).unsqueeze(2)
    target = (
        torch.sum(term1 * w1, dim=1) + torch.sum(term2 * w2, dim=1)
        + torch.sum(term3 * w1, dim=1) + terminal
    )                                                              # method.py:675-690
    learned = -torch.einsum("ij,...j->...i", st.sigma.t(), gv)
    wanted = -torch.einsum("ij,...j->...i", st.sigma.t(), target)
    if use_stopping_time:                                          # method.py:692-720
        mask = stop_ind.unsqueeze(2)
        learned, wanted = mask * learned, mask * wanted
        norm = torch.sum(stop_ind)
    else:
        norm = K1 * B
    obj = torch.sum((learned - wanted) ** 2 * weight.unsqueeze(0).unsqueeze(2)) / norm
    return obj, torch.mean(weight), torch.std(weight)              # method.py:897-906


# --------------------------------------------------------------------------------------
# Kink-free test inputs
# --------------------------------------------------------------------------------------
def unet_preactivations64(unet: Dict[str, Tensor], tx: Tensor):
    """The six pre-ReLU tensors of FullyConnectedUNet.forward (models.py:233-242) in fp64."""
    P = {k: v.double() for k, v in unet.items()}

    def lin(name: str, v: Tensor) -> Tensor:
        return F.linear(v, P[name + ".0.weight"], P[name + ".0.bias"])

    z1 = lin("down_0", tx)
    r1 = torch.relu(z1)
    z2 = lin("down_1", r1)
    r2 = torch.relu(z2)
    z3 = lin("down_2", r2)
    r3 = torch.relu(z3)
    y2 = lin("up_2", r3)
    o2 = torch.relu(y2) + lin("res_2", r2)
    y1 = lin("up_1", o2)
    o1 = torch.relu(y1) + lin("res_1", r1)
    return [z1, z2, z3, y2, y1, lin("up_0", o1)]


def path_noise(seed: int, attempts, K: int, d: int) -> Tensor:
    """(K, B, d) standard normal increments, path m drawn from numpy's default_rng([seed, m, attempts[m]]): a fixture
    stores ``seed`` and the small integer array ``attempts`` instead of the noise itself."""
    import numpy as np

    cols = [np.random.default_rng([int(seed), m, int(a)]).standard_normal((K, d)).astype(np.float32)
            for m, a in enumerate(attempts)]
    return torch.from_numpy(np.stack(cols, axis=1))


def kink_free_attempts(st: Setting, unet: Dict[str, Tensor], x0: Tensor, ts: Tensor, seed: int, B: int,
                       warm: Optional[WarmStartTable] = None, rel_delta: float = 4e-6, max_rounds: int = 400):
    """Redraw the Brownian increments of every path that comes within ``rel_delta * rms(layer)`` of a ReLU kink of
    the control network at any of its K+1 points (fp64 evaluation along the oracle's rollout) until none does.

    d loss / d theta is discontinuous where a pre-activation crosses zero: the mask of that unit is decided by the
    last bits of a 64..256-term sum, so any two fp32 evaluations of the same network (MKL vs cuBLAS, FFMA vs tensor
    cores) may pick different masks there, and ONE flipped unit moves a gradient tensor of a few-thousand-point batch
    by ~1e-3 although both results are valid sub-gradients.  On paths that keep a safe distance from every kink the
    comparison measures arithmetic.  Returns the integer array ``attempts`` for :func:`path_noise`."""
    import numpy as np

    K, d = ts.shape[0] - 1, st.d
    attempts = np.zeros(B, dtype=np.int64)
    for _ in range(max_rounds):
        noises = path_noise(seed, attempts, K, d)
        states = rollout(st, unet, x0.repeat(B, 1), ts, noises=noises, warm=warm)[0]
        tx = torch.cat([ts.reshape(-1, 1, 1).expand(K + 1, B, 1), states], -1).double()
        near = torch.zeros(B, dtype=torch.bool)
        for z in unet_preactivations64(unet, tx):
            near |= (z.abs() < rel_delta * z.pow(2).mean().sqrt()).any(-1).any(0)
        if not bool(near.any()):
            return attempts
        attempts[near.numpy()] += 1
    raise RuntimeError("kink_free_attempts: no kink-free draw found")


# --------------------------------------------------------------------------------------
# Philox4x32-10 (counter-based RNG used by the CUDA rollout when noise is not injected)
# --------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """numpy uint32 Philox4x32-10 (Salmon et al. 2011, the published algorithm); checks
    the in-kernel generator bit for bit.  counter: (..., 4) uint32, key: (2,) uint32."""
    import numpy as np

    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c = [counter[..., i].astype(np.uint32) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c[0].astype(np.uint64) * M0
            p1 = c[2].astype(np.uint64) * M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return np.stack(c, axis=-1)
