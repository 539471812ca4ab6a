"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/<case>.npz

The reference (facebookresearch/SOC-matching) is imported from where it lies, with the
three import stubs SURVEY.md section 8c lists (nvidia_smi, omegaconf, tqdm.notebook).
No reference source is copied.  For every case the file stores the inputs (setting
constants, network parameters, x0, ts, injected noise) and the reference's outputs:
the 8-tuple of ``utils.stochastic_trajectories`` (utils.py:17-128) and, for the loss
cases, ``SOC_Solver.loss(...)[0]`` (method.py:223-906), ``mean(w)``, ``std(w)`` and
``autograd.grad`` of the objective w.r.t. every UNet / M-net / gamma parameter.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = os.environ.get("SOCM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    """Import SOC_matching from /root/reference with the three stubs."""
    if "nvidia_smi" not in sys.modules:
        sys.modules["nvidia_smi"] = types.ModuleType("nvidia_smi")
    if "omegaconf" not in sys.modules:
        class _AttrDict(dict):
            __getattr__ = dict.__getitem__

        oc = types.ModuleType("omegaconf")
        oc.OmegaConf = SimpleNamespace(create=lambda d: _AttrDict(d))
        sys.modules["omegaconf"] = oc
    try:
        import tqdm.notebook  # noqa: F401
    except Exception:
        tn = types.ModuleType("tqdm.notebook")
        tn.trange = range
        sys.modules["tqdm.notebook"] = tn
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from SOC_matching import method, models, utils  # noqa: F401
    from SOC_matching.experiment_settings import (  # noqa: F401
        OU_linear, OU_quadratic, double_well, molecular_dynamics,
    )
    utils.trange = range
    return SimpleNamespace(
        method=method, models=models, utils=utils, OU_Quadratic=OU_quadratic.OU_Quadratic,
        OU_Linear=OU_linear.OU_Linear, DoubleWell=double_well.DoubleWell,
        MolecularDynamics=molecular_dynamics.MolecularDynamics,
    )


class _ReplayNoise:
    """Context manager: torch.randn_like pops the next slice of a fixed (K,B,d) tensor,
    so that SOC_Solver.loss re-uses a known noise (utils.py:39 is the only call site)."""

    def __init__(self, noises):
        self.noises, self.k = noises, 0

    def __enter__(self):
        self.orig = torch.randn_like

        def fake(x, *a, **kw):
            out = self.noises[self.k].clone()
            assert out.shape == x.shape
            self.k += 1
            return out

        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        torch.randn_like = self.orig


def seeded_params(module, seed, scale=1.0):
    """Overwrite every parameter with numpy-seeded uniform(-1,1)/sqrt(fan_in)*scale.
    (Used for the full-size case so that the fixture stores a seed, not 0.8 MB of
    weights; tests/ regenerate them with the same rule, see tests/helpers.py.)"""
    rng = np.random.default_rng(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            fan_in = p.shape[1] if p.dim() == 2 else p.shape[0]
            if p.dim() == 1 and name.endswith("bias"):
                # bias of Linear(in, out): bound by the matching weight's fan_in
                w = dict(module.named_parameters())[name[:-4] + "weight"]
                fan_in = w.shape[1]
            vals = rng.uniform(-1.0, 1.0, size=tuple(p.shape)).astype(np.float32)
            p.copy_(torch.from_numpy(vals) * (scale / np.sqrt(fan_in)))


def build_case(ref, name):
    """Returns (sde, solver_kwargs, case dict)."""
    c = CASES[name]
    torch.manual_seed(c["seed"])
    d, K, B = c["d"], c["K"], c["B"]
    kind = c["kind"]
    common = dict(device="cpu", dim=d, hdims=c["hdims"], hdims_M=c["hdims_M"], lmbd=c["lmbd"],
                  gamma=c["gamma"], scaling_factor_nabla_V=c.get("sf_v", 1.0),
                  scaling_factor_M=c.get("sf_m", 0.1))
    tens = {}
    if kind == "ou_quadratic":
        # settings.py:215-228 (easy: 0.2/0.2/0.1, hard: 1/1/0.5); a non-diagonal variant
        # is used here on purpose so that index conventions are exercised.
        x0 = 0.5 * torch.randn(d)
        sigma = torch.eye(d) + (0.1 * torch.randn(d, d) if c.get("dense") else 0)
        a, p, q = c["apq"]
        pert = (lambda: 0.05 * torch.randn(d, d)) if c.get("dense") else (lambda: 0)
        A = a * torch.eye(d) + pert()
        P = p * torch.eye(d) + pert()
        Q = q * torch.eye(d) + pert()
        sde = ref.OU_Quadratic(A=A, P=P, Q=Q, sigma=sigma, **common)
        tens.update(A=A, P=P, Q=Q)
    elif kind == "ou_linear":
        # settings.py:237-243
        x0 = torch.zeros(d)
        xi = 0.1 * torch.randn(d, d)
        omega = torch.ones(d)
        A = -torch.eye(d) + xi
        sigma = torch.eye(d) + xi
        sde = ref.OU_Linear(A=A, omega=omega, sigma=sigma, **common)
        tens.update(A=A, omega=omega)
    elif kind == "double_well":
        # settings.py:252-267
        x0 = torch.zeros(d)
        kappa, nu = torch.ones(d), torch.ones(d)
        kappa[:3], nu[:3] = 5, 3
        sigma = torch.eye(d)
        sde = ref.DoubleWell(kappa=kappa, nu=nu, sigma=sigma, **common)
        tens.update(kappa=kappa, nu=nu)
    elif kind == "molecular_dynamics":
        # settings.py:269-289
        x0 = -torch.ones(d)
        kappa = torch.ones(d)
        sigma = torch.eye(d)
        # use_stopping_time selects the M-network class (method.py:117-132): TwoBoundarySigmoidMLP or SigmoidMLP
        sde = ref.MolecularDynamics(kappa=kappa, sigma=sigma, use_stopping_time=bool(c.get("stopping")), **common)
        tens.update(kappa=kappa)
    else:
        raise ValueError(kind)
    sde.initialize_models()
    if c.get("param_seed") is not None:
        seeded_params(sde.nabla_V, c["param_seed"], c.get("sf_v", 1.0))
        seeded_params(sde.M.sigmoid_layers, c["param_seed"] + 1, c.get("sf_m", 0.1))
    tens.update(x0=x0, sigma=sigma)
    return sde, tens


def fit_warm_start(ref, sde, x0, sigma, K, n_iter, lr):
    """settings.py:117-140 with a short spline fit (cost per call is independent of
    fit quality); returns the reference RestrictedControl."""
    cfg = SimpleNamespace(
        optim=SimpleNamespace(splines_lr=lr),
        method=SimpleNamespace(num_iterations_splines=n_iter),
    )
    # the fit uses problem.b/f/g/sigma/T of a ground-truth SDE object; the neural SDE has them too
    res = ref.utils.restricted_SOC(sde, x0.unsqueeze(0), torch.zeros(1, x0.shape[0]), "cpu", cfg)
    return ref.models.RestrictedControl(res["gpath"], sigma, sde.b, "cpu", 1.0, 1)


def probe_affine(gpath, t_shift, d):
    """gpath.ut(t', x) is affine in x (gsbm_lib.py:227-306): recover (A, c) with d+1 probes."""
    pts = torch.cat([torch.zeros(1, d), torch.eye(d)], 0)            # (d+1, d)
    x = pts[None, :, None, :]                                        # (1, N, 1, d)
    with torch.no_grad():
        out = gpath.ut(t_shift.reshape(1), x, direction="fwd", create_graph_jvp=False)[0, :, 0, :]
    c = out[0]
    A = (out[1:] - c).t().contiguous()                               # column j = f(e_j) - f(0)
    return A, c


def warm_table(ws, ts, T=1.0):
    d = ws.sigma.shape[0]
    K = ts.shape[0] - 1
    A_roll, c_roll, A_loss, c_loss = [], [], [], []
    for k in range(K + 1):
        t = ts[k]
        if k < K:   # rank-2 branch, models.py:169-170
            tt = torch.tensor([t])
            tt = tt + 1e-4 if tt < T / 2 else tt - 1e-4
            A, c = probe_affine(ws.gpath, tt, d)
            A_roll.append(A), c_roll.append(c)
        tl = t.clone()  # rank-3 branch, models.py:184-186
        if t < T / 2:
            tl = t + 1e-4
        elif t > T / 2:
            tl = t - 1e-4
        if k == 0 or k == K:
            pass
        A, c = probe_affine(ws.gpath, tl.reshape(1), d)
        A_loss.append(A), c_loss.append(c)
    return torch.stack(A_roll), torch.stack(c_roll), torch.stack(A_loss), torch.stack(c_loss)


CASES = {
    # name: reduced-size versions of the five BASELINE.json configs (+ one full-width net)
    "c1_ou_quadratic_easy": dict(kind="ou_quadratic", d=4, K=12, B=6, hdims=[24, 16, 8], hdims_M=[12, 12],
                                 lmbd=1.0, gamma=2.0, apq=(0.2, 0.2, 0.1), seed=0,
                                 algorithms=["SOCM", "SOCM_const_M", "SOCM_exp", "SOCM_adjoint", "cross_entropy", "log-variance", "variance", "moment"]),
    "c1b_ou_quadratic_dense": dict(kind="ou_quadratic", d=3, K=9, B=5, hdims=[16, 12, 8], hdims_M=[10, 10],
                                   lmbd=0.7, gamma=1.5, apq=(0.3, 0.4, 0.2), dense=True, seed=1,
                                   algorithms=["SOCM", "SOCM_const_M", "SOCM_exp", "SOCM_adjoint", "cross_entropy", "log-variance", "variance", "moment"]),
    "c2_ou_linear": dict(kind="ou_linear", d=3, K=10, B=5, hdims=[24, 16, 8], hdims_M=[12, 12],
                         lmbd=1.0, gamma=2.0, seed=2, algorithms=["SOCM", "SOCM_const_M", "SOCM_exp", "SOCM_adjoint"]),
    "c3_ou_quadratic_hard_warm": dict(kind="ou_quadratic", d=3, K=10, B=5, hdims=[24, 16, 8], hdims_M=[12, 12],
                                      lmbd=1.0, gamma=2.0, apq=(1.0, 1.0, 0.5), sf_v=0.1, seed=3,
                                      warm=dict(n_iter=4, lr=2e-4), algorithms=["SOCM", "SOCM_const_M"]),
    # SOCM_const_M / SOCM_exp / SOCM_adjoint never read use_stopping_time (method.py:289-478, 722-749): plain dts,
    # no mask, normaliser (K+1) B, although the rollout stops paths (utils.py:33 keys on Phi alone)
    "c4_molecular_dynamics": dict(kind="molecular_dynamics", d=1, K=60, B=32, hdims=[24, 16, 8], hdims_M=[8, 8],
                                  lmbd=1.0, gamma=2.0, seed=8, stopping=True,
                                  algorithms=["SOCM", "cross_entropy", "log-variance", "moment", "SOCM_const_M",
                                              "SOCM_adjoint"]),
    # the same setting run WITHOUT use_stopping_time: SigmoidMLP M(t, s), plain dts in the SOCM target, no mask
    "c4b_molecular_dynamics_plain": dict(kind="molecular_dynamics", d=1, K=60, B=32, hdims=[24, 16, 8],
                                         hdims_M=[8, 8], lmbd=1.0, gamma=2.0, seed=8, stopping=False,
                                         algorithms=["SOCM", "SOCM_const_M", "SOCM_exp", "cross_entropy",
                                                     "log-variance"]),
    "c5_double_well": dict(kind="double_well", d=4, K=100, B=5, hdims=[24, 16, 8], hdims_M=[12, 12],
                           lmbd=1.0, gamma=6.0, seed=5,
                           algorithms=["SOCM", "SOCM_const_M", "SOCM_exp", "SOCM_adjoint", "cross_entropy", "log-variance", "variance", "moment"]),
    "c5_double_well_fullnet": dict(kind="double_well", d=10, K=100, B=3, hdims=[256, 128, 64], hdims_M=[128, 128],
                                   lmbd=1.0, gamma=6.0, seed=6, param_seed=1234, algorithms=["SOCM", "SOCM_adjoint"]),
    # BASELINE config 3 at the DEFAULT width: the warm-start table enters the tcgen05 / FFMA-tile kernels (the
    # reduced-width c3 above only reaches the shape-generic ones).  The reference evaluates its own RestrictedControl
    # (two jvp's per step); the product and the oracle the tabulated affine form.
    "c3_full_warm": dict(kind="ou_quadratic", d=20, K=30, B=24, hdims=[256, 128, 64], hdims_M=[128, 128],
                         lmbd=1.0, gamma=2.0, apq=(1.0, 1.0, 0.5), sf_v=0.1, seed=13, param_seed=2468,
                         warm=dict(n_iter=4, lr=2e-4), algorithms=["SOCM", "SOCM_const_M"]),
    # default width at (K+1) B = 65 536 trajectory points = SOCM_LOSS_TC_MIN_POINTS: the default dispatch runs the
    # tcgen05 K3, so this fixture pins the tensor-core path against the reference's OWN loss and gradients.  The
    # injected noise is regenerated from `noise_seed` (numpy) and only a subsample of the trajectories is stored.
    "big_c5_double_well_tc": dict(kind="double_well", d=10, K=63, B=1024, hdims=[256, 128, 64], hdims_M=[128, 128],
                                  lmbd=1.0, gamma=6.0, seed=16, param_seed=4321, noise_seed=77, keep_paths=48,
                                  algorithms=["SOCM"]),
}


def run_case(ref, name):
    c = CASES[name]
    sde, tens = build_case(ref, name)
    d, K, B = c["d"], c["K"], c["B"]
    x0, sigma = tens["x0"], tens["sigma"]
    ts = torch.linspace(0, 1.0, K + 1)
    out = {f"setting/{k}": v.numpy() for k, v in tens.items()}
    meta = dict(kind=c["kind"], d=d, K=K, B=B, lmbd=c["lmbd"], gamma=c["gamma"], hdims=c["hdims"],
                hdims_M=c["hdims_M"], stopping=bool(c.get("stopping")), warm=bool(c.get("warm")),
                algorithms=c["algorithms"], param_seed=c.get("param_seed"), sf_v=c.get("sf_v", 1.0),
                sf_m=c.get("sf_m", 0.1))
    ws = None
    if c.get("warm"):
        ws = fit_warm_start(ref, sde, x0, sigma, K, **c["warm"])
        sde.u_warm_start, sde.use_warm_start = ws, True
        A_r, c_r, A_l, c_l = warm_table(ws, ts)
        out.update({"warm/A_roll": A_r.numpy(), "warm/c_roll": c_r.numpy(),
                    "warm/A_loss": A_l.numpy(), "warm/c_loss": c_l.numpy()})
    if c.get("param_seed") is None:
        for k, v in sde.nabla_V.state_dict().items():
            out[f"unet/{k}"] = v.numpy().copy()
        for k, v in sde.M.state_dict().items():
            out[f"mnet/{k}"] = v.numpy().copy()
    out["ts"] = ts.numpy()

    # ---- rollout (utils.py:17-128) with the reference's own RNG; keep the noise it drew
    torch.manual_seed(1000 + c["seed"])
    state0 = x0.repeat(B, 1)
    names = ["states", "noises", "stop_indicators", "fractional_timesteps", "logw_det", "logw_sto",
             "logw_term", "controls"]
    if c.get("noise_seed") is not None:      # big case: numpy-seeded noise, replayed through torch.randn_like
        # the noise is selected (NOT the reference's computation on it): paths that keep clear of the ReLU kinks of
        # the control network, where the gradient is well defined (oracle/socm_oracle.py:kink_free_attempts)
        root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
        sys.path[:0] = [root, os.path.join(root, "tests")]
        from oracle import socm_oracle as orc
        from helpers import seeded_unet          # the same seeded parameters seeded_params() wrote into the reference
        st = orc.Setting(c["kind"], d, sigma, c["lmbd"], **{k: v for k, v in tens.items() if k not in ("x0", "sigma")})
        attempts = orc.kink_free_attempts(st, seeded_unet(d, c["hdims"], c["param_seed"], c.get("sf_v", 1.0)), x0, ts,
                                          c["noise_seed"], B)
        print(f"  {name}: {int((attempts > 0).sum())} of {B} paths redrawn to stay clear of ReLU kinks")
        meta["noise_seed"], meta["keep_paths"] = c["noise_seed"], c["keep_paths"]
        out["noise_attempts"] = attempts.astype(np.int16)
        fixed = orc.path_noise(c["noise_seed"], attempts, K, d)
        with _ReplayNoise(fixed):
            traj = ref.utils.stochastic_trajectories(sde, state0, ts, c["lmbd"])
        kp = c["keep_paths"]
        for n, v in zip(names, traj):
            v = v.float()
            if n == "noises":
                continue
            out[f"rollout_sub/{n}"] = (v[:, :kp] if v.dim() >= 2 else v).numpy().copy()
    else:
        traj = ref.utils.stochastic_trajectories(sde, state0, ts, c["lmbd"])
        for n, v in zip(names, traj):
            out[f"rollout/{n}"] = v.float().numpy().copy()
    noises = traj[1]
    if c.get("stopping"):
        n_stopped = int((traj[2][-1] == 0).sum())
        print(f"  {name}: {n_stopped}/{B} paths stopped")
        meta["n_stopped"] = n_stopped

    # ---- loss + grads (method.py:223-906) replaying the same noise
    solver = ref.method.SOC_Solver(sde, x0, None, T=1.0, num_steps=K, lmbd=c["lmbd"], d=d, sigma=sigma)
    for algo in c["algorithms"]:
        params = list(sde.nabla_V.named_parameters())
        pm = [(f"sigmoid_layers.{n}", p) for n, p in sde.M.sigmoid_layers.named_parameters()]
        gam = [("gamma", sde.gamma)]
        if c.get("stopping"):
            gam += [("gamma2", sde.gamma2), ("gamma3", sde.gamma3)]
        if algo == "SOCM_exp":   # main.py:166-169: the decay rate of M_t(s) = exp(-gamma (s - t)) I lives on the solver
            solver.gamma = torch.nn.Parameter(torch.tensor([float(c["gamma"])]))
        with _ReplayNoise(noises):
            res = solver.loss(B, algorithm=algo, u_warm_start=ws, use_warm_start=bool(ws),
                              use_stopping_time=bool(c.get("stopping")))
        obj = res[0]
        wanted = (params + (pm + gam if algo == "SOCM" else []) + ([("gamma", solver.gamma)] if algo == "SOCM_exp" else [])
                  + ([("y0", solver.y0)] if algo == "moment" else []))
        grads = torch.autograd.grad(obj, [p for _, p in wanted], allow_unused=True)
        out[f"{algo}/loss"] = obj.detach().numpy()
        out[f"{algo}/weight_mean"] = res[5].detach().numpy()
        out[f"{algo}/weight_std"] = res[6].detach().numpy()
        for (n, p), g in zip(wanted, grads):
            grp = "unet" if any(n is q for q, _ in params) else ("gam" if n.startswith(("gamma", "y0")) else "mnet")
            out[f"{algo}/grad/{grp}/{n}"] = (torch.zeros_like(p) if g is None else g).numpy()
        print(f"  {name} [{algo}]: loss={float(obj):.6g} mean_w={float(res[5]):.4g}")
    out["solver_y0"] = solver.y0.detach().numpy().copy()
    out["gammas"] = np.array([float(sde.gamma), float(getattr(sde, "gamma2", 1.0)), float(getattr(sde, "gamma3", 1.0))],
                             dtype=np.float32)
    import json
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def main(argv):
    torch.set_num_threads(1)  # bit-stable reductions
    ref = import_reference()
    for name in (argv or list(CASES)):
        print(name)
        run_case(ref, name)


if __name__ == "__main__":
    main(sys.argv[1:])
