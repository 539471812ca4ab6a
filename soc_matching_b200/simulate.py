"""Drop-in for the reference's ``utils.stochastic_trajectories`` (utils.py:17-128) and the two
Monte-Carlo evaluators that loop it (utils.py:131-231), on the fused CUDA rollout kernel."""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import _lib, networks
from .sde import SettingDesc, describe_setting

_SEED_COUNTER = [0]
# tensor-core engine of the default-width kernels: None = library default (SOCM_F16 environment switch),
# "f16" = fp16-split engine with two CTAs per SM (csrc/unet_h.cuh, d <= 15), "tf32" = 3xTF32 engine (csrc/unet_tc.cuh)
ENGINE = None


def sync_engine(lib):
    """Hand ENGINE to the library for the calls that have no engine flag of their own (the target GEMMs)."""
    _lib.check(lib.socm_set_default_engine({None: -1, "tf32": 0, "f16": 1}[ENGINE]))


def step_table(t: torch.Tensor, lmbd: float) -> torch.Tensor:
    """[5][K] fp32: dt_k, sqrt(lmbd dt_k), dt_k/lmbd, sqrt(dt_k/lmbd), t_k -- the per-step scalars of
    utils.py:38,47,95-98 computed with the same fp32 torch ops (dt from the fp32 linspace)."""
    t = t.detach().float()
    dt = t[1:] - t[:-1]
    return torch.stack([dt, torch.sqrt(lmbd * dt), dt / lmbd, torch.sqrt(dt / lmbd), t[:-1]]).contiguous()


def next_seed() -> int:
    """Philox key for one rollout call: torch's global seed (main.py:71 seeds it once) mixed with a
    call counter, so successive calls draw fresh, reproducible noise."""
    _SEED_COUNTER[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _SEED_COUNTER[0]) & 0xFFFFFFFFFFFFFFFF


def resolve_warm_start(sde, t: torch.Tensor):
    """WarmStartTable of ``sde`` on grid ``t`` (None if the sde runs without warm start)."""
    ws = getattr(sde, "u_warm_start", None)
    if not (getattr(sde, "use_warm_start", False) and ws):
        return None
    if isinstance(ws, networks.WarmStartTable):
        return ws
    if hasattr(ws, "gpath"):  # a reference RestrictedControl: tabulate once per grid and cache
        key = (t.shape[0], float(t[0]), float(t[-1]))
        cache = getattr(ws, "_socm_table_cache", None)
        if cache is None or cache[0] != key:
            table = networks.WarmStartTable.from_restricted_control(ws, t, getattr(ws, "T", 1.0)).to(t.device)
            ws._socm_table_cache = (key, table)
        return ws._socm_table_cache[1]
    raise NotImplementedError(f"warm-start object of type {type(ws).__name__} cannot be tabulated")


def _warm_struct(A: torch.Tensor, c: torch.Tensor):
    w = _lib.WarmTable()
    w.A, w.c = A.data_ptr(), c.data_ptr()
    return w


class RolloutWorkspace:
    """Reusable device buffers for one (B, K, d) shape."""

    def __init__(self, desc: SettingDesc, net, B: int, K: int, device, store_traj: bool = True):
        d = desc.d
        f32 = dict(device=device, dtype=torch.float32)
        self.B, self.K = B, K
        if store_traj:
            self.states = torch.empty(K + 1, B, d, **f32)
            self.noises = torch.empty(K, B, d, **f32)
            self.controls = torch.empty(K, B, d, **f32)
            self.stop = torch.empty(K + 1, B, **f32)
            self.eff_dt = torch.empty(K, B, **f32)
        else:
            self.states = self.noises = self.controls = self.stop = self.eff_dt = None
        self.lw = torch.empty(3, B, **f32)
        self.packed = None
        if net is not None:      # packed weight tape of the control network (none under a tabulated control)
            udesc, keep = networks.unet_desc(net)
            nbytes = _lib.load().socm_rollout_workspace_bytes(udesc)
            self.packed = torch.empty((nbytes + 3) // 4, **f32)
            del keep


def rollout(sde, x0: torch.Tensor, t: torch.Tensor, lmbd: float, *, noises: Optional[torch.Tensor] = None,
            seed: Optional[int] = None, path_offset: int = 0, store_traj: bool = True,
            force_generic: bool = False, force_ffma: bool = False, desc: Optional[SettingDesc] = None,
            workspace: Optional[RolloutWorkspace] = None, timer=None) -> RolloutWorkspace:
    """Run K1 once.  Returns the workspace holding the outputs.  Kernel selection (csrc/rollout.cu):
    default hdims -> tcgen05 tensor-core kernel (3xTF32); ``force_ffma`` -> fp32 FFMA tile kernel;
    ``force_generic`` / other hdims -> shape-generic warp-per-path kernel."""
    lib = _lib.load()
    sync_engine(lib)
    _lib.require_cuda(x0, "x0")
    desc = desc or describe_setting(sde, x0.device)
    B, K = int(x0.shape[0]), int(t.shape[0]) - 1
    if desc.lmbd != float(lmbd):
        desc.c_struct.lmbd = float(lmbd)
    if not getattr(sde, "use_learned_control", False):
        return _rollout_tabulated(lib, sde, desc, x0, t, lmbd, noises, seed, path_offset, store_traj, workspace, timer)
    ws = workspace or RolloutWorkspace(desc, sde.nabla_V, B, K, x0.device, store_traj)
    udesc, keep = networks.unet_desc(sde.nabla_V)
    tab = step_table(t.to(x0.device), float(lmbd))
    x0c = x0.detach().float().contiguous()
    warm = resolve_warm_start(sde, t)
    wstruct = _warm_struct(warm.A_roll, warm.c_roll) if warm is not None else None
    flags = ((0 if store_traj else _lib.ROLLOUT_NO_TRAJ) | (_lib.ROLLOUT_FORCE_GENERIC if force_generic else 0)
             | (_lib.ROLLOUT_FORCE_FFMA if force_ffma else 0)
             | {None: 0, "f16": _lib.ROLLOUT_F16, "tf32": _lib.ROLLOUT_TF32}[ENGINE])
    noise_ptr = None
    if noises is not None:
        _lib.require_cuda(noises, "noises")
        noises = noises.detach().float().contiguous()
        assert tuple(noises.shape) == (K, B, desc.d), f"noises must be (K,B,d), got {tuple(noises.shape)}"
        noise_ptr = noises.data_ptr()
        if store_traj:
            ws.noises = noises
    if seed is None:
        seed = next_seed()
    if B == 0:  # empty batch: nothing to launch (empty tensors have no device pointer)
        return ws
    args = (desc.c_struct, udesc, wstruct, _lib.ptr(x0c), _lib.ptr(tab), noise_ptr, seed, path_offset, B, K,
            _lib.ptr(ws.states), _lib.ptr(ws.noises), _lib.ptr(ws.controls), _lib.ptr(ws.stop), _lib.ptr(ws.eff_dt),
            ws.lw[0].data_ptr(), ws.lw[1].data_ptr(), ws.lw[2].data_ptr(), _lib.ptr(ws.packed), flags,
            _lib.stream_ptr())
    if timer is not None:
        # kernels of one call: tcgen05 path = fold + pack + rollout, FFMA tile path = pack + rollout, generic = 1
        default_arch = (udesc.h0, udesc.h1, udesc.h2) == (256, 128, 64)
        n_k = 1 if (force_generic or not default_arch) else (2 if (force_ffma or udesc.d > 23) else 3)
        if n_k == 3 and udesc.d <= 15 and ENGINE != "tf32" and os.environ.get("SOCM_F16") != "0":
            n_k = 4   # fp16-split engine: fold + calibration + pack + rollout
        timer("rollout", n_k, lib.socm_rollout_f32, *args)
    else:
        _lib.check(lib.socm_rollout_f32(*args))
    del keep
    return ws


def _rollout_tabulated(lib, sde, desc, x0, t, lmbd, noises, seed, path_offset, store_traj, workspace, timer):
    """method.py:103-107: ``sde.u`` is a tabulated ground-truth control (models.py:10-150), no network in the loop."""
    from . import controls
    if getattr(sde, "u", None) is None:
        raise NotImplementedError("sde.use_learned_control is False and sde.u is None: there is no control to roll out")
    B, K, d = int(x0.shape[0]), int(t.shape[0]) - 1, desc.d
    ws = workspace or RolloutWorkspace(desc, None, B, K, x0.device, store_traj)
    ctrl, keep = controls.tabulate(sde.u, t, d, x0.device)
    tab = step_table(t.to(x0.device), float(lmbd))
    x0c = x0.detach().float().contiguous()
    noise_ptr = None
    if noises is not None:
        _lib.require_cuda(noises, "noises")
        noises = noises.detach().float().contiguous()
        assert tuple(noises.shape) == (K, B, d), f"noises must be (K,B,d), got {tuple(noises.shape)}"
        noise_ptr = noises.data_ptr()
        if store_traj:
            ws.noises = noises
    if B == 0:
        return ws
    args = (desc.c_struct, ctrl, _lib.ptr(x0c), _lib.ptr(tab), noise_ptr, next_seed() if seed is None else seed,
            path_offset, B, K, _lib.ptr(ws.states), _lib.ptr(ws.noises), _lib.ptr(ws.controls), _lib.ptr(ws.stop),
            _lib.ptr(ws.eff_dt), ws.lw[0].data_ptr(), ws.lw[1].data_ptr(), ws.lw[2].data_ptr(),
            0 if store_traj else _lib.ROLLOUT_NO_TRAJ, _lib.stream_ptr())
    if timer is not None:
        timer("rollout", 1, lib.socm_rollout_tabulated_f32, *args)
    else:
        _lib.check(lib.socm_rollout_tabulated_f32(*args))
    del keep
    return ws


def stochastic_trajectories(sde, x0, t, lmbd, detach=True, verbose=False, *, noises=None, seed=None,
                            force_generic=False, force_ffma=False):
    """Same signature and 8-tuple as utils.py:17-128:
    (states (K+1,B,d), noises (K,B,d), stop_indicators (K+1,B), fractional_timesteps (K,B),
     log_path_weight_deterministic (B,), log_path_weight_stochastic (B,), log_terminal_weight (B,),
     controls (K,B,d)).

    Extensions (keyword-only): ``noises`` injects the Brownian increments (parity tests), ``seed``
    fixes the Philox key.  ``detach=False`` (only used by the reference's rel_entropy loss, out of
    scope) is not supported: the fused kernel does not record an autograd graph."""
    if not detach:
        raise NotImplementedError("detach=False (back-propagation through the rollout) is not supported")
    ws = rollout(sde, x0, t, lmbd, noises=noises, seed=seed, force_generic=force_generic, force_ffma=force_ffma)
    return (ws.states, ws.noises, ws.stop, ws.eff_dt, ws.lw[0], ws.lw[1], ws.lw[2], ws.controls)


def control_objective(sde, x0, ts, lmbd, batch_size, total_n_samples=65536, verbose=False):
    """utils.py:131-163 on the weights-only rollout (no trajectory is stored)."""
    n_batches = int(total_n_samples // batch_size)
    n = n_batches * batch_size
    state0 = x0.repeat(n, 1)
    ws = rollout(sde, state0, ts.to(state0), lmbd, store_traj=False)
    losses = -lmbd * (ws.lw[0] + ws.lw[2])
    return torch.mean(losses), torch.std(losses) / math.sqrt(n - 1)


def normalization_constant(sde, x0, ts, cfg, n_batches_normalization=512, ground_truth_control=None):
    """utils.py:166-231: Monte-Carlo estimate of E[exp(log-weights)] under the current control (and, if a
    ground-truth control is given, the importance-weighted squared control error), from
    ``n_batches_normalization`` rollouts of the batch ``x0``.  Without a ground-truth control the rollouts
    run in weights-only mode (no trajectory leaves the SM); they are grouped into launches of at most
    2^17 paths.  ``cfg.method.lmbd`` is the only field of ``cfg`` that is read, as in the reference."""
    lmbd = float(cfg.method.lmbd)
    B, d = int(x0.shape[0]), int(x0.shape[1])
    ts = ts.to(x0)
    K = int(ts.shape[0]) - 1
    per_launch = max(1, (1 << 17) // max(B, 1))
    store = ground_truth_control is not None
    logw, sqd = [], torch.zeros((), device=x0.device, dtype=torch.float64)
    for k0 in range(0, n_batches_normalization, per_launch):
        nb = min(per_launch, n_batches_normalization - k0)
        ws = rollout(sde, x0.repeat(nb, 1), ts, lmbd, store_traj=store)
        lw = (ws.lw[0] + ws.lw[1] + ws.lw[2]).reshape(nb, B)
        logw.append(lw)
        if store:
            gt = ground_truth_control(ts, ws.states, t_is_tensor=True)[:-1].detach()
            sqd += torch.sum(((gt - ws.controls) ** 2).double() * torch.exp(lw).double().reshape(1, -1, 1)) / (K * B)
    log_weights = torch.cat(logw, dim=0).t()                # (B, n_batches), as torch.stack(..., dim=1)
    weights = torch.exp(log_weights)
    print(f"Average and std. dev. of log_weights for all batches: {torch.mean(log_weights)} {torch.std(log_weights)}")
    n = weights.shape[0] * weights.shape[1]
    norm_sqd_diff_mean = (sqd / n_batches_normalization).float() if store else None
    return torch.mean(weights), torch.std(weights) / math.sqrt(n - 1), norm_sqd_diff_mean
