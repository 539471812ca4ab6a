"""Network containers of the hot path.

Same parameter names and shapes as the reference's modules (models.py:202-393) so that
``state_dict``s move freely between the two; the arithmetic of the control network runs in
the CUDA kernels (csrc/), these modules only hold the parameters and hand raw pointers to
the C ABI.  The small, B-independent M-networks are evaluated with torch ops on the GPU
(they feed the block-triangular table L once per iteration, see mtable.py).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from . import _lib

UNET_LAYER_ORDER = ("down_0", "down_1", "down_2", "res_0", "res_1", "res_2", "up_2", "up_1", "up_0")


def _scaled_linear(n_in: int, n_out: int, scale: float) -> nn.Linear:
    lin = nn.Linear(n_in, n_out)
    with torch.no_grad():
        lin.weight.mul_(scale)
        lin.bias.mul_(scale)
    return lin


class FullyConnectedUNet(nn.Module):
    """Control network nabla_V (reference: models.py:202-242).

    r1 = relu(down_0 [t,x]), r2 = relu(down_1 r1), r3 = relu(down_2 r2),
    o2 = relu(up_2 r3) + res_2 r2, o1 = relu(up_1 o2) + res_1 r1, out = relu(up_0 o1) + res_0 [t,x].
    """

    def __init__(self, dim: int = 2, hdims: Sequence[int] = (256, 128, 64), scaling_factor: float = 1.0):
        super().__init__()
        h0, h1, h2 = hdims
        self.dim, self.hdims = dim, tuple(hdims)
        shapes = {
            "down_0": (dim + 1, h0), "down_1": (h0, h1), "down_2": (h1, h2),
            "res_0": (dim + 1, dim), "res_1": (h0, h0), "res_2": (h1, h1),
            "up_2": (h2, h1), "up_1": (h1, h0), "up_0": (h0, dim),
        }
        for name in UNET_LAYER_ORDER:  # registration order == reference's named_parameters order
            n_in, n_out = shapes[name]
            layers: List[nn.Module] = [_scaled_linear(n_in, n_out, scaling_factor)]
            if not name.startswith("res"):
                layers.append(nn.ReLU())
            setattr(self, name, nn.Sequential(*layers))

    def linears(self) -> List[nn.Linear]:
        return [getattr(self, n)[0] for n in UNET_LAYER_ORDER]

    def forward(self, tx: torch.Tensor) -> torch.Tensor:
        """nabla_V at n points, (n, d+1) -> (n, d), through socm_unet_forward_f32 (no autograd:
        gradients of the control network are produced by the fused loss kernel)."""
        _lib.require_cuda(tx, "tx")
        flat = tx.detach().reshape(-1, self.dim + 1).contiguous().float()
        out = torch.empty(flat.shape[0], self.dim, device=tx.device, dtype=torch.float32)
        desc, keep = unet_desc(self)
        _lib.check(_lib.load().socm_unet_forward_f32(desc, _lib.ptr(flat), flat.shape[0], _lib.ptr(out),
                                                     _lib.stream_ptr()))
        del keep
        return out.reshape(*tx.shape[:-1], self.dim)


def unet_desc(net: nn.Module):
    """(ctypes socm_unet, keep-alive list) from any module with the reference's layer names."""
    lins = [getattr(net, n)[0] for n in UNET_LAYER_ORDER]
    d = lins[3].out_features
    h0, h1, h2 = lins[0].out_features, lins[1].out_features, lins[2].out_features
    desc = _lib.UNet()
    desc.d, desc.h0, desc.h1, desc.h2 = d, h0, h1, h2
    keep = []
    for i, lin in enumerate(lins):
        w, b = lin.weight.detach(), lin.bias.detach()
        _lib.require_cuda(w, "UNet weights")
        if not w.is_contiguous():
            w = w.contiguous()
        if w.dtype != torch.float32:
            raise _lib.SocmError("UNet parameters must be fp32")
        keep += [w, b]
        desc.w[i] = w.data_ptr()
        desc.b[i] = b.data_ptr()
    return desc, keep


def unet_parameters(net: nn.Module) -> List[nn.Parameter]:
    """Parameters in the flat order the kernels use for gradients: (w, b) per layer."""
    out = []
    for n in UNET_LAYER_ORDER:
        lin = getattr(net, n)[0]
        out += [lin.weight, lin.bias]
    return out


class SigmoidMLP(nn.Module):
    """M(t,s) = e^{-gamma(s-t)} I + (1 - e^{-gamma(s-t)}) N(t,s)   (reference: models.py:245-275)."""

    def __init__(self, dim: int = 10, hdims: Sequence[int] = (128, 128), gamma=3.0, scaling_factor: float = 1.0):
        super().__init__()
        self.dim, self.gamma, self.scaling_factor = dim, gamma, scaling_factor
        self.sigmoid_layers = nn.Sequential(
            _scaled_linear(2, hdims[0], scaling_factor), nn.ReLU(),
            _scaled_linear(hdims[0], hdims[1], scaling_factor), nn.ReLU(),
            _scaled_linear(hdims[1], dim * dim, scaling_factor),
        )

    def forward(self, t: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
        return self.value_and_ds(t, s)[0]

    def value_and_ds(self, t: torch.Tensor, s: torch.Tensor):
        """(M, dM/ds) for P pairs, each (P, d, d).  dM/ds by the analytic forward-mode tangent
        of the ReLU MLP instead of the reference's d^2 reverse passes
        (functorch.jacrev, method.py:510-515); both are differentiable w.r.t. parameters."""
        d = self.dim
        l1, l2, l3 = self.sigmoid_layers[0], self.sigmoid_layers[2], self.sigmoid_layers[4]
        ts = torch.stack((t, s), dim=1)
        z1 = torch.addmm(l1.bias, ts, l1.weight.t())
        h1 = torch.relu(z1)
        z2 = torch.addmm(l2.bias, h1, l2.weight.t())
        h2 = torch.relu(z2)
        net = torch.addmm(l3.bias, h2, l3.weight.t()).reshape(-1, d, d)
        dh1 = (z1 > 0).to(ts.dtype) * l1.weight[:, 1].unsqueeze(0)
        dh2 = (z2 > 0).to(ts.dtype) * (dh1 @ l2.weight.t())
        dnet = (dh2 @ l3.weight.t()).reshape(-1, d, d)
        decay = (1.0 / torch.exp(self.gamma * (s - t))).reshape(-1, 1, 1)
        eye = torch.eye(d, device=ts.device, dtype=ts.dtype).unsqueeze(0)
        m = decay * eye + (1 - decay) * net
        dm = self.gamma * decay * (net - eye) + (1 - decay) * dnet
        return m, dm


class TwoBoundarySigmoidMLP(nn.Module):
    """Stopping-time aware M(t, s, tau)   (reference: models.py:278-393)."""

    def __init__(self, dim=10, hdims=(128, 128), gamma=3.0, gamma2=3.0, gamma3=3.0, scaling_factor=1.0, T=1.0):
        super().__init__()
        self.dim, self.gamma, self.gamma2, self.gamma3, self.T = dim, gamma, gamma2, gamma3, T
        self.scaling_factor = scaling_factor
        self.sigmoid_layers = nn.Sequential(
            _scaled_linear(3, hdims[0], scaling_factor), nn.ReLU(),
            _scaled_linear(hdims[0], hdims[1], scaling_factor), nn.ReLU(),
            _scaled_linear(hdims[1], dim * dim, scaling_factor),
        )

    def forward(self, t: torch.Tensor, s: torch.Tensor, tau: torch.Tensor) -> torch.Tensor:
        """t, s: (P,), tau: (P, B)  ->  (P, B, d, d)."""
        d, dev = self.dim, t.device
        col = lambda v: v.unsqueeze(1)  # noqa: E731
        stopped_in = torch.cat((col(t), col(s), col(torch.zeros_like(s))), dim=1)
        alive_in = torch.cat((col(t), col(s), col(torch.ones_like(s))), dim=1)
        net_stopped = self.sigmoid_layers(stopped_in).reshape(-1, 1, d, d)
        net_alive = self.sigmoid_layers(alive_in).reshape(-1, 1, d, d)
        eye = torch.eye(d, device=dev).reshape(1, 1, d, d)
        one = torch.ones(1, device=dev)
        ratio = col(1 - torch.exp(-self.gamma * (s - t))) / (
            1 - torch.exp(-self.gamma * torch.abs(tau - col(t))) + 1e-7
        )
        fac = torch.nan_to_num(1 - torch.minimum(ratio, one), nan=0.0)
        fac = fac * (tau - 1e-3 > col(s)).to(torch.int)
        decay3 = col(torch.exp(-self.gamma3 * (s - t)))
        alive = (tau > self.T - 1e-3).to(torch.int)
        ident_part = ((1 - alive) * fac + alive * decay3).unsqueeze(2).unsqueeze(3) * eye
        g2 = self.gamma2
        bump = (1 - torch.exp(-g2 * fac)) * (torch.exp(-g2 * fac) - torch.exp(-g2))
        net_part = ((1 - alive) * bump).unsqueeze(2).unsqueeze(3) * net_stopped + (
            alive * (1 - decay3)
        ).unsqueeze(2).unsqueeze(3) * net_alive
        return ident_part + net_part

    def value_and_ds(self, t, s, tau):
        """(M, nan_to_num(dM/ds)); forward-mode tangent in s (each pair depends on its own s only)."""
        m, dm = torch.func.jvp(lambda sv: self.forward(t, sv, tau), (s,), (torch.ones_like(s),))
        return m, torch.nan_to_num(dm)


class WarmStartTable:
    """Warm-start control u_ws(t_k, x) = sigma^{-1}(c_k + A_k x - b(x)) tabulated on the time grid
    (reference: models.RestrictedControl, models.py:153-199, on a frozen Gaussian-path spline).

    ``A_roll/c_roll`` (K rows) use the rank-2 branch's time shift (models.py:170),
    ``A_loss/c_loss`` (K+1 rows) the rank-3 branch's (models.py:184-186)."""

    def __init__(self, A_roll, c_roll, A_loss, c_loss):
        self.A_roll, self.c_roll = A_roll.contiguous().float(), c_roll.contiguous().float()
        self.A_loss, self.c_loss = A_loss.contiguous().float(), c_loss.contiguous().float()

    def to(self, device):
        return WarmStartTable(self.A_roll.to(device), self.c_roll.to(device), self.A_loss.to(device),
                              self.c_loss.to(device))

    def __bool__(self):  # the reference tests `if self.use_warm_start and self.u_warm_start`
        return True

    @staticmethod
    def _probe(gpath, t_shift: torch.Tensor, d: int):
        pts = torch.cat([torch.zeros(1, d), torch.eye(d)], 0).to(t_shift.device)
        with torch.no_grad():
            out = gpath.ut(t_shift.reshape(1), pts[None, :, None, :], direction="fwd",
                           create_graph_jvp=False)[0, :, 0, :]
        c = out[0]
        return (out[1:] - c).t().contiguous(), c

    @classmethod
    def from_restricted_control(cls, ws, ts: torch.Tensor, T: float = 1.0):
        """Tabulate a reference ``RestrictedControl`` (anything with ``.gpath.ut`` affine in x and
        ``.sigma``) at the grid times: d+1 probes per time."""
        d = ws.sigma.shape[0]
        K = ts.shape[0] - 1
        Ar, cr, Al, cl = [], [], [], []
        for k in range(K + 1):
            t = ts[k].detach().cpu()
            if k < K:
                tt = t + 1e-4 if t < T / 2 else t - 1e-4
                A, c = cls._probe(ws.gpath, tt, d)
                Ar.append(A), cr.append(c)
            tl = t + 1e-4 if t < T / 2 else (t - 1e-4 if t > T / 2 else t)
            A, c = cls._probe(ws.gpath, tl, d)
            Al.append(A), cl.append(c)
        return cls(torch.stack(Ar), torch.stack(cr), torch.stack(Al), torch.stack(cl))
