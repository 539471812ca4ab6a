"""Shared test helpers: golden-fixture loading, seeded parameter generation, error norms."""
from __future__ import annotations

import glob
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

from oracle import socm_oracle as orc  # noqa: E402  (tests are allowed to use the oracle)

UNET_LAYERS = ["down_0", "down_1", "down_2", "res_0", "res_1", "res_2", "up_2", "up_1", "up_0"]


def golden_names(big=False):
    """Fixture names; the ``big_*`` fixtures (subsampled trajectories, seeded noise) are listed separately."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n.startswith("big_") == big]


def rel_l2(a, b) -> float:
    """||a-b||_2 / ||b||_2 (norm-wise); 0 if both are exactly zero."""
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    den = float(torch.linalg.norm(b))
    num = float(torch.linalg.norm(a - b))
    if den == 0.0:
        return 0.0 if num == 0.0 else float("inf")
    return num / den


def unet_shapes(d, hdims):
    h0, h1, h2 = hdims
    return {
        "down_0": (h0, d + 1), "down_1": (h1, h0), "down_2": (h2, h1),
        "res_0": (d, d + 1), "res_1": (h0, h0), "res_2": (h1, h1),
        "up_2": (h1, h2), "up_1": (h0, h1), "up_0": (d, h0),
    }


def mnet_shapes(d, hdims_M, n_in=2):
    return {"0": (hdims_M[0], n_in), "2": (hdims_M[1], hdims_M[0]), "4": (d * d, hdims_M[1])}


def _seeded(named_shapes, seed, scale):
    """numpy-seeded uniform(-1,1)/sqrt(fan_in)*scale, weights then bias per layer, in the
    reference's ``named_parameters`` order (must match oracle/make_golden.py:seeded_params)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in named_shapes:
        fan_in = shape[1]
        w = rng.uniform(-1.0, 1.0, size=shape).astype(np.float32)
        out[name + ".weight"] = torch.from_numpy(w) * (scale / np.sqrt(fan_in))
        b = rng.uniform(-1.0, 1.0, size=(shape[0],)).astype(np.float32)
        out[name + ".bias"] = torch.from_numpy(b) * (scale / np.sqrt(fan_in))
    return out


def seeded_unet(d, hdims, seed, scale=1.0):
    shp = unet_shapes(d, hdims)
    return _seeded([(f"{n}.0", shp[n]) for n in UNET_LAYERS], seed, scale)


def seeded_mnet(d, hdims_M, seed, scale=0.1, n_in=2):
    shp = mnet_shapes(d, hdims_M, n_in)
    return _seeded([(f"sigmoid_layers.{k}", shp[k]) for k in ("0", "2", "4")], seed, scale)


class Golden:
    """One tests/golden/<name>.npz written by oracle/make_golden.py."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.name = name
        self.meta = json.loads(bytes(z["meta_json"]).decode())
        m = self.meta
        t = lambda k: torch.from_numpy(z[k].copy())  # noqa: E731
        kw = {}
        for k in ("A", "P", "Q", "omega", "kappa", "nu"):
            if f"setting/{k}" in z.files:
                kw[k] = t(f"setting/{k}")
        self.setting = orc.Setting(kind=m["kind"], d=m["d"], sigma=t("setting/sigma"), lmbd=m["lmbd"], **kw)
        self.x0 = t("setting/x0")
        self.ts = t("ts")
        if m.get("param_seed") is not None:
            self.unet = seeded_unet(m["d"], m["hdims"], m["param_seed"], m["sf_v"])
            self.mnet = seeded_mnet(m["d"], m["hdims_M"], m["param_seed"] + 1, m["sf_m"],
                                    3 if m["stopping"] else 2)
        else:
            self.unet = {k[5:]: t(k) for k in z.files if k.startswith("unet/")}
            self.mnet = {k[5:]: t(k) for k in z.files if k.startswith("mnet/")}
        g = z["gammas"]
        self.gammas = {"gamma": torch.tensor([float(g[0])]), "gamma2": torch.tensor([float(g[1])]),
                       "gamma3": torch.tensor([float(g[2])])}
        if "solver_y0" in z.files:   # SOC_Solver.y0 (method.py:172), used by the "moment" loss only
            self.gammas["y0"] = t("solver_y0").reshape(1)
        self.warm = None
        if m["warm"]:
            self.warm = orc.WarmStartTable(t("warm/A_roll"), t("warm/c_roll"), t("warm/A_loss"), t("warm/c_loss"))
        self.rollout_names = ["states", "noises", "stop_indicators", "fractional_timesteps",
                              "logw_det", "logw_sto", "logw_term", "controls"]
        self.big = m.get("noise_seed") is not None
        if self.big:
            # full noise regenerated from the seed (+ per-path redraw counts, kink-free selection); the reference's
            # trajectories are stored for the first
            # `keep_paths` paths only (log-weights for all paths): self.traj_sub, self.noises
            self.noises = orc.path_noise(m["noise_seed"], z["noise_attempts"], m["K"], m["d"])
            self.traj_sub = {n: t(f"rollout_sub/{n}") for n in self.rollout_names if n != "noises"}
            self.traj = None
        else:
            self.traj = tuple(t(f"rollout/{n}") for n in self.rollout_names)
            self.noises = self.traj[1]

    def grads(self, algo):
        pre = f"{algo}/grad/"
        return {k[len(pre):]: torch.from_numpy(self.z[k].copy()) for k in self.z.files if k.startswith(pre)}

    def scalar(self, key):
        return float(self.z[key])


# ---------------------------------------------------------------------------------------------
# product-side builders (GPU tests)
def make_product_sde(setting: "orc.Setting", unet: dict, mnet: dict, gammas: dict, hdims, hdims_M, device,
                     stopping=False, warm=None):
    """soc_matching_b200 SDE object carrying the given constants / parameters."""
    import soc_matching_b200 as sb
    st = setting
    dv = lambda t: None if t is None else t.to(device)  # noqa: E731
    common = dict(device=device, dim=st.d, hdims=hdims, hdims_M=hdims_M, lmbd=st.lmbd, sigma=dv(st.sigma),
                  gamma=float(gammas["gamma"]))
    if st.kind == "ou_quadratic":
        sde = sb.OU_Quadratic(A=dv(st.A), P=dv(st.P), Q=dv(st.Q), **common)
    elif st.kind == "ou_linear":
        sde = sb.OU_Linear(A=dv(st.A), omega=dv(st.omega), **common)
    elif st.kind == "double_well":
        sde = sb.DoubleWell(kappa=dv(st.kappa), nu=dv(st.nu), **common)
    else:
        sde = sb.MolecularDynamics(kappa=dv(st.kappa), use_stopping_time=stopping, gamma2=float(gammas["gamma2"]),
                                   gamma3=float(gammas["gamma3"]), **common)
    sde.initialize_models()
    sde.nabla_V.load_state_dict({k: v.to(device) for k, v in unet.items()})
    sde.M.sigmoid_layers.load_state_dict(
        {k[len("sigmoid_layers."):]: v.to(device) for k, v in mnet.items() if k.startswith("sigmoid_layers")})
    if warm is not None:
        sde.u_warm_start = sb.WarmStartTable(warm.A_roll, warm.c_roll, warm.A_loss, warm.c_loss).to(device)
        sde.use_warm_start = True
    return sde


def random_setting(kind: str, d: int, seed: int, lmbd: float = 1.0, dense_sigma: bool = False) -> "orc.Setting":
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    eye = torch.eye(d)
    sigma = eye + 0.1 * rn(d, d) if dense_sigma else eye.clone()
    if kind == "ou_quadratic":
        return orc.Setting(kind, d, sigma, lmbd, A=0.2 * eye + 0.05 * rn(d, d), P=0.2 * eye + 0.05 * rn(d, d),
                           Q=0.1 * eye + 0.05 * rn(d, d))
    if kind == "ou_linear":
        xi = 0.1 * rn(d, d)
        return orc.Setting(kind, d, eye + xi, lmbd, A=-eye + xi, omega=torch.ones(d))
    kappa, nu = torch.ones(d), torch.ones(d)
    kappa[:3], nu[:3] = 5, 3
    if kind == "double_well":
        return orc.Setting(kind, d, sigma, lmbd, kappa=kappa, nu=nu)
    return orc.Setting(kind, d, eye.clone(), lmbd, kappa=torch.ones(d))


def gpu_kink_free_noise(sde, x0, ts, B, seed, rel_delta=4e-6, max_rounds=60, slab=16):
    """GPU twin of ``orc.kink_free_attempts`` for batches the CPU oracle cannot roll out (a full bench chunk): injected
    noise (K, B, d) on the device such that no path comes within ``rel_delta * rms(layer)`` of a ReLU kink of the control
    network -- rollout by the product's exact-fp32 FFMA kernel, pre-activations by torch fp64 on the GPU in slabs of
    grid times.  See the oracle function for why gradient parity is only well defined on such paths."""
    import soc_matching_b200 as sb  # noqa: F401
    from soc_matching_b200 import simulate
    dev = x0.device
    K, d = ts.shape[0] - 1, sde.dim
    gen = torch.Generator(device=dev).manual_seed(seed)
    noises = torch.randn(K, B, d, device=dev, generator=gen)
    unet = {k: v.detach() for k, v in sde.nabla_V.state_dict().items()}
    x0b = x0.reshape(1, d).expand(B, d).contiguous()
    rms = None
    for _ in range(max_rounds):
        states = simulate.rollout(sde, x0b, ts, sde.lmbd, noises=noises, force_ffma=True).states
        near = torch.zeros(B, dtype=torch.bool, device=dev)
        sums = [0.0] * 6
        for i0 in range(0, K + 1, slab):
            i1 = min(K + 1, i0 + slab)
            tx = torch.cat([ts[i0:i1].reshape(-1, 1, 1).expand(i1 - i0, B, 1), states[i0:i1]], -1).double()
            pre = orc.unet_preactivations64(unet, tx)
            if rms is None:
                for j, z in enumerate(pre):
                    sums[j] += float(z.pow(2).sum())
            else:
                for j, z in enumerate(pre):
                    near |= (z.abs() < rel_delta * rms[j]).any(-1).any(0)
        if rms is None:   # first pass only measures the layer scales (they barely move when a few paths are redrawn)
            widths = [256, 128, 64, 128, 256, d]
            rms = [math.sqrt(s / ((K + 1) * B * w)) for s, w in zip(sums, widths)]
            continue
        n = int(near.sum())
        if n == 0:
            return noises
        noises[:, near] = torch.randn(K, n, d, device=dev, generator=gen)
    raise AssertionError("gpu_kink_free_noise: no kink-free draw found")
