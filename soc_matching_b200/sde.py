"""Problem objects: the NeuralSDE interface of the reference (method.py:15-143) and the four
benchmark settings (experiment_settings/*.py) with the same constructor arguments, attribute
names and closed-form primitives.  The primitives here are thin torch expressions kept for
API compatibility (callers such as evaluation code may use them); the hot path never calls
them -- it passes the setting's constants to the CUDA kernels through ``describe_setting``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, networks

KIND_IDS = {"OU_Quadratic": 0, "OU_Linear": 1, "DoubleWell": 2, "MolecularDynamics": 3}


class NeuralSDE(nn.Module):
    """dX = (b(X) + sigma u(t,X)) dt + sqrt(lmbd) sigma dW with u = -sigma^T nabla_V (+ warm start)."""

    noise_type = "diagonal"
    sde_type = "ito"

    def __init__(self, device="cuda", dim=2, hdims=(256, 128, 64), hdims_M=(128, 128), u=None, lmbd=1.0,
                 sigma=None, gamma=1.0, gamma2=1.0, gamma3=1.0, scaling_factor_nabla_V=1.0,
                 scaling_factor_M=1.0, T=1.0, u_warm_start=None, use_warm_start=False,
                 use_stopping_time=False):
        super().__init__()
        self.device, self.dim = device, dim
        self.hdims, self.hdims_M = list(hdims), list(hdims_M)
        self.u, self.lmbd = u, lmbd
        self.sigma = torch.eye(dim, device=device) if sigma is None else sigma
        self.gamma, self.gamma2, self.gamma3 = gamma, gamma2, gamma3
        self.scaling_factor_nabla_V, self.scaling_factor_M = scaling_factor_nabla_V, scaling_factor_M
        self.use_learned_control = False
        self.T = T
        self.u_warm_start, self.use_warm_start = u_warm_start, use_warm_start
        self.use_stopping_time = use_stopping_time

    def initialize_models(self):
        """method.py:109-143: creates nabla_V, M and the gamma parameters (shared with M)."""
        dev = self.device
        self.nabla_V = networks.FullyConnectedUNet(self.dim, self.hdims, self.scaling_factor_nabla_V).to(dev)
        self.gamma = nn.Parameter(torch.tensor([float(self.gamma)], device=dev))
        if self.use_stopping_time:
            self.gamma2 = nn.Parameter(torch.tensor([float(self.gamma2)], device=dev))
            self.gamma3 = nn.Parameter(torch.tensor([float(self.gamma3)], device=dev))
            self.M = networks.TwoBoundarySigmoidMLP(self.dim, self.hdims_M, self.gamma, self.gamma2, self.gamma3,
                                                    self.scaling_factor_M).to(dev)
        else:
            self.M = networks.SigmoidMLP(self.dim, self.hdims_M, self.gamma, self.scaling_factor_M).to(dev)
        self.use_learned_control = True

    def control(self, t, x, verbose=False):
        """method.py:58-107.  Learned control through the CUDA UNet kernel (no autograd)."""
        if not self.use_learned_control:
            return None if self.u is None else self.u(t, x)
        if x.dim() == 2:
            tx = torch.cat([t.reshape(-1, 1).expand(x.shape[0], 1), x], dim=-1)
        else:
            tx = torch.cat([t.reshape(-1, 1, 1).expand(x.shape[0], x.shape[1], 1), x], dim=-1)
        u = -torch.einsum("ij,...j->...i", self.sigma.t(), self.nabla_V(tx))
        if self.use_warm_start and self.u_warm_start:
            raise _lib.SocmError("control() with a warm start is only available inside the fused rollout")
        return u

    # closed-form primitives, overridden by the settings
    def b(self, t, x):
        raise NotImplementedError

    def nabla_b(self, t, x):
        raise NotImplementedError

    def f(self, t, x):
        raise NotImplementedError

    def nabla_f(self, t, x):
        raise NotImplementedError

    def g(self, x):
        raise NotImplementedError

    def nabla_g(self, x):
        raise NotImplementedError


def _kw(kwargs, **defaults):
    out = dict(defaults)
    out.update(kwargs)
    return out


class OU_Quadratic(NeuralSDE):
    """b = Ax, f = x'Px, g = x'Qx   (OU_quadratic.py:9-83)."""

    def __init__(self, A=None, P=None, Q=None, **kw):
        super().__init__(**_kw(kw, gamma=kw.get("gamma", 3.0)))
        self.A, self.P, self.Q = A, P, Q

    def b(self, t, x):
        return x @ self.A.t()

    def nabla_b(self, t, x):
        return self.A.t().expand(*x.shape[:-1], *self.A.shape)

    def f(self, t, x):
        return (x * (x @ self.P.t())).sum(-1)

    def nabla_f(self, t, x):
        return 2 * (x @ self.P.t())

    def g(self, x):
        return (x * (x @ self.Q.t())).sum(-1)

    def nabla_g(self, x):
        return 2 * (x @ self.Q.t())


class OU_Linear(NeuralSDE):
    """b = Ax, f = 0, g = omega . x   (OU_linear.py:9-96)."""

    def __init__(self, A=None, omega=None, **kw):
        super().__init__(**_kw(kw, gamma=kw.get("gamma", 3.0)))
        self.A, self.omega = A, omega

    def b(self, t, x):
        return x @ self.A.t()

    def nabla_b(self, t, x):
        return self.A.t().expand(*x.shape[:-1], *self.A.shape)

    def f(self, t, x):
        return torch.zeros_like(x[..., 0])

    def nabla_f(self, t, x):
        return torch.zeros_like(x)

    def g(self, x):
        return x @ self.omega

    def nabla_g(self, x):
        return self.omega.expand_as(x).clone()


class DoubleWell(NeuralSDE):
    """b_i = -4 kappa_i x_i (x_i^2 - 1), f = 0, g = sum nu_i (x_i^2-1)^2   (double_well.py:12-97)."""

    def __init__(self, kappa=None, nu=None, **kw):
        super().__init__(**_kw(kw, gamma=kw.get("gamma", 3.0)))
        self.kappa, self.nu = kappa, nu

    def b(self, t, x):
        return -2 * self.kappa * (x**2 - 1) * 2 * x

    def nabla_b(self, t, x):
        return -torch.diag_embed(8 * self.kappa * x**2 + 4 * self.kappa * (x**2 - 1))

    def f(self, t, x):
        return torch.zeros_like(x[..., 0])

    def nabla_f(self, t, x):
        return torch.zeros_like(x)

    def g(self, x):
        return (self.nu * (x**2 - 1) ** 2).sum(-1)

    def nabla_g(self, x):
        return 2 * self.nu * (x**2 - 1) * 2 * x


class MolecularDynamics(NeuralSDE):
    """Double-well drift, f = 1, g = 0, stops when Phi(x) = -x_0 <= 0   (molecular_dynamics.py:11-95)."""

    def __init__(self, kappa=None, **kw):
        super().__init__(**_kw(kw, gamma=kw.get("gamma", 3.0)))
        self.kappa = kappa

    def b(self, t, x):
        return -2 * self.kappa * (x**2 - 1) * 2 * x

    def nabla_b(self, t, x):
        return -torch.diag_embed(8 * self.kappa * x**2 + 4 * self.kappa * (x**2 - 1))

    def f(self, t, x):
        return torch.ones_like(x[..., 0])

    def nabla_f(self, t, x):
        return torch.zeros_like(x)

    def g(self, x):
        return torch.zeros_like(x[..., 0])

    def nabla_g(self, x):
        return torch.zeros_like(x)

    def Phi(self, x):
        return -x[..., 0]


# --------------------------------------------------------------------------------------------
@dataclass
class SettingDesc:
    """Device-resident constants of a setting + the ctypes struct handed to the kernels."""

    kind: int
    d: int
    lmbd: float
    tensors: dict            # keeps the fp32 CUDA tensors alive
    c_struct: _lib.Setting
    has_stopping: bool


def describe_setting(sde, device: Optional[torch.device] = None) -> SettingDesc:
    """Extract the closed-form constants of ``sde`` (this package's classes or the reference's,
    matched by class name: an unknown subclass raises, there is no generic fallback)."""
    name = type(sde).__name__
    if name not in KIND_IDS:
        raise NotImplementedError(
            f"setting {name!r} is not one of {sorted(KIND_IDS)}; the fused kernels need its closed-form "
            "drift/cost and there is no fallback path by design"
        )
    kind = KIND_IDS[name]
    d = int(sde.dim)
    sigma = sde.sigma.detach()
    dev = torch.device(device) if device is not None else sigma.device
    if dev.type != "cuda":
        raise _lib.SocmError(f"setting tensors live on {dev}; soc_matching_b200 needs CUDA")

    def dv(t):
        return t.detach().to(device=dev, dtype=torch.float32).contiguous()

    tens = {"sigma": dv(sigma), "sigma_inv": dv(torch.inverse(sigma.float()))}
    need = {0: ("A", "P", "Q"), 1: ("A", "omega"), 2: ("kappa", "nu"), 3: ("kappa",)}[kind]
    for k in need:
        tens[k] = dv(getattr(sde, k))
    cs = _lib.Setting()
    cs.kind, cs.d, cs.lmbd = kind, d, float(sde.lmbd)
    cs.sigma_is_identity = int(bool(torch.equal(tens["sigma"], torch.eye(d, device=dev))))
    for k in ("sigma", "sigma_inv", "A", "P", "Q", "omega", "kappa", "nu"):
        setattr(cs, k, tens[k].data_ptr() if k in tens else None)
    return SettingDesc(kind, d, float(sde.lmbd), tens, cs, hasattr(sde, "Phi"))


def make_benchmark_sde(setting: str, d: int, device="cuda", hdims=(256, 128, 64), hdims_M=(128, 128), lmbd=1.0,
                       gamma=2.0, scaling_factor_nabla_V=1.0, scaling_factor_M=0.1, use_stopping_time=False):
    """The benchmark problems exactly as experiment_settings/settings.py:210-289 builds them
    (constants only; ground-truth controls and the warm-start fit are callers' business).
    Returns (x0, sigma, neural_sde)."""
    common = dict(device=device, dim=d, hdims=hdims, hdims_M=hdims_M, lmbd=lmbd, gamma=gamma,
                  scaling_factor_nabla_V=scaling_factor_nabla_V, scaling_factor_M=scaling_factor_M)
    eye = torch.eye(d, device=device)
    if setting in ("OU_quadratic_easy", "OU_quadratic_hard"):
        x0 = torch.tensor([0.4, 0.6], device=device) if d == 2 else 0.5 * torch.randn(d).to(device)
        a, p, q = (1.0, 1.0, 0.5) if setting.endswith("hard") else (0.2, 0.2, 0.1)
        sde = OU_Quadratic(A=a * eye, P=p * eye, Q=q * eye, sigma=eye.clone(), **common)
    elif setting == "OU_linear":
        x0 = torch.zeros(d, device=device)
        xi = 0.1 * torch.randn(d, d).to(device)
        sde = OU_Linear(A=-eye + xi, omega=torch.ones(d, device=device), sigma=eye + xi, **common)
    elif setting == "double_well":
        x0 = torch.zeros(d, device=device)
        kappa, nu = torch.ones(d, device=device), torch.ones(d, device=device)
        kappa[:3], nu[:3] = 5, 3
        sde = DoubleWell(kappa=kappa, nu=nu, sigma=eye.clone(), **common)
    elif setting == "molecular_dynamics":
        x0 = -torch.ones(d, device=device)
        sde = MolecularDynamics(kappa=torch.ones(d, device=device), sigma=eye.clone(),
                                use_stopping_time=use_stopping_time, **common)
    else:
        raise NotImplementedError(f"unknown setting {setting!r}")
    sde.initialize_models()
    return x0, sde.sigma, sde
