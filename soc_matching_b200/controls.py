"""Tabulated controls: the ground-truth / baseline controls of the reference (models.py:10-150) that an SDE object
carries in ``sde.u`` when ``use_learned_control`` is False (method.py:103-107).  ``control_objective`` and
``normalization_constant`` run the rollout under them on the optimal SDE (utils.py:131-231, main.py:117-153), and the
L2-error metric evaluates them on whole trajectories (``t_is_tensor=True``, method.py:858-875).

The classes keep the reference's constructor arguments, attribute names and call signature ``u(t, x,
t_is_tensor=False)``; the fused rollout does not call them -- ``tabulate`` turns them (or the reference's own objects,
matched by class name) into the per-grid-time tables of ``socm_tab_control`` with the reference's index rules.
"""
from __future__ import annotations

import torch

from . import _lib


def _grid_index(n_rows: int, t: torch.Tensor, T: float) -> torch.Tensor:
    return torch.floor(n_rows * t / T).to(torch.int64)


class LinearControl:
    """u(t, x) = U[floor((n-1) t / T)] x with U of shape (n, d, d)   (models.py:10-39)."""

    def __init__(self, u, T):
        self.u, self.T = u, T

    def evaluate(self, t):
        return self.u[_grid_index(self.u.shape[0] - 1, t, self.T)]

    evaluate_tensor = evaluate

    def __call__(self, t, x, t_is_tensor=False):
        mat = self.evaluate(t)
        if not t_is_tensor:
            mat = mat if mat.dim() == 2 else mat[0]
            return x @ mat.t()
        spec = "bij,bj->bi" if x.dim() == 2 else "aij,abj->abi"
        return torch.einsum(spec, mat, x)


class ConstantControlLinear:
    """u(t) = ut[floor(n t / T)] (state independent), ut of shape (n, d)   (models.py:61-81; the t_is_tensor branch
    indexes with n - 1 instead of n, kept as in the reference)."""

    def __init__(self, ut, T):
        self.ut, self.T = ut, T

    def evaluate_ut(self, t):
        return self.ut[_grid_index(self.ut.shape[0], t, self.T)]

    def evaluate_ut_tensor(self, t):
        return self.ut[_grid_index(self.ut.shape[0] - 1, t, self.T), :]

    def __call__(self, t, x, t_is_tensor=False):
        if not t_is_tensor:
            return self.evaluate_ut(t).unsqueeze(0).repeat(x.shape[0], 1)
        return self.evaluate_ut_tensor(t).unsqueeze(1).repeat(1, x.shape[1], 1)


class LowDimControl:
    """Coordinate-wise lookup u_j(t, x) = ut[ceil(t / delta_t), clamp(floor((x_j + xb) / delta_x)), j] with ut of shape
    (nt, nx, d): the finite-difference solution of the double-well problem   (models.py:84-150)."""

    def __init__(self, ut, T, xb, dim, delta_t, delta_x):
        self.ut, self.T, self.xb, self.dim, self.delta_t, self.delta_x = ut, T, xb, dim, delta_t, delta_x

    def _lookup(self, t_col, x_flat):
        it = torch.ceil(t_col / self.delta_t).to(torch.int64)
        ix = torch.floor((x_flat + self.xb) / self.delta_x).to(torch.int64).clamp(0, self.ut.shape[1] - 1)
        j = torch.arange(self.dim, device=x_flat.device).unsqueeze(0).expand_as(ix)
        return self.ut[it.unsqueeze(1).expand_as(ix), ix, j]

    def __call__(self, t, x, t_is_tensor=False):
        if not t_is_tensor:
            flat = x.reshape(-1, self.dim)
            t_col = torch.as_tensor(t, device=x.device, dtype=flat.dtype).reshape(-1).expand(flat.shape[0])
            return self._lookup(t_col, flat)
        t_col = t.reshape(-1, 1).expand(x.shape[0], x.shape[1]).reshape(-1)
        return self._lookup(t_col, x.reshape(-1, self.dim)).reshape(x.shape)


def tabulate(u, ts: torch.Tensor, d: int, device):
    """``socm_tab_control`` (+ the tensors it points to) of control object ``u`` on the grid times ``ts[:-1]``.
    Objects are matched by class name, so the reference's own LinearControl / ConstantControlLinear / LowDimControl
    work unchanged; anything else raises (no fallback by design)."""
    name = type(u).__name__
    t = ts[:-1].detach().to(device=device, dtype=torch.float32)
    K = t.shape[0]
    ctrl, keep = _lib.TabControl(), []

    def dv(x, dtype=torch.float32):
        x = x.detach().to(device=device, dtype=dtype).contiguous()
        keep.append(x)
        return x

    if name == "LinearControl":
        tab = u.u.to(device)
        A = dv(tab[_grid_index(tab.shape[0] - 1, t, u.T)])                  # models.py:15-18
        assert tuple(A.shape) == (K, d, d), f"LinearControl table must be (n, d, d), got {tuple(tab.shape)}"
        ctrl.kind, ctrl.A = _lib.CONTROL_AFFINE, A.data_ptr()
    elif name == "ConstantControlLinear":
        tab = u.ut.to(device)
        c = dv(tab[_grid_index(tab.shape[0], t, u.T)])                      # models.py:66-69 (non-tensor branch)
        assert tuple(c.shape) == (K, d), f"ConstantControlLinear table must be (n, d), got {tuple(tab.shape)}"
        ctrl.kind, ctrl.c = _lib.CONTROL_AFFINE, c.data_ptr()
    elif name == "LowDimControl":
        tab = dv(u.ut)
        assert tab.dim() == 3 and tab.shape[2] == d, f"LowDimControl table must be (nt, nx, d), got {tuple(tab.shape)}"
        it = torch.ceil(t / u.delta_t).to(torch.int64)                       # models.py:97
        if int(it.max()) >= tab.shape[0]:
            raise _lib.SocmError("LowDimControl: the time grid runs past the table")
        it32 = dv(it, torch.int32)
        ctrl.kind, ctrl.ut, ctrl.idx_t = _lib.CONTROL_LOOKUP, tab.data_ptr(), it32.data_ptr()
        ctrl.nx, ctrl.xb, ctrl.dx = int(tab.shape[1]), float(u.xb), float(u.delta_x)
    else:
        raise NotImplementedError(
            f"control of type {name!r}: the fused rollout runs the learned control (UNet) or the tabulated "
            "LinearControl / ConstantControlLinear / LowDimControl (models.py:10-150)")
    return ctrl, keep
