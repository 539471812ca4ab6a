"""The training step around the hot path: what the body of the reference's loop does
between ``soc_solver.loss(...)`` and the next iteration (main.py:298-393) -- divide by the running normalisation
constant, ``backward()``, gradient statistics, Adam step + ``zero_grad``, EMA bookkeeping -- with the statistics in
one CUDA launch (csrc/ema.cu) and the optimiser in another (csrc/adam.cu).  Nothing here synchronises with the host:
every statistic stays a device scalar until the caller reads it (the reference prints every 10th iteration,
main.py:415-423)."""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _lib
from .optim import FusedAdam

STAT_NAMES = ("grad_norm_sqd", "EMA_grad_norm_sqd", "sqd_norm_EMA_grad", "EMA_loss", "EMA_weight_mean",
              "EMA_weight_std", "normalization_const")


class TrainingStatistics:
    """EMA state of main.py:325-393 on the device.  ``update`` is one kernel launch."""

    def __init__(self, grad_params: Iterable[torch.nn.Parameter], normalization_const, ema_coeff: float = 0.01,
                 ema_weight_mean_coeff: float = 0.002):
        self.params = list(grad_params)                       # main.py:327: soc_solver.neural_sde.nabla_V.parameters()
        dev = self.params[0].device
        _lib.require_cuda(self.params[0], "parameters")
        self.ema_coeff, self.ema_weight_mean_coeff = float(ema_coeff), float(ema_weight_mean_coeff)
        self.ema_grad = [torch.zeros_like(p) for p in self.params]
        self.stats = torch.zeros(8, device=dev, dtype=torch.float32)
        self.stats[6] = float(normalization_const)            # Monte-Carlo estimate of main.py:117-137
        self._scratch = torch.zeros(4, device=dev, dtype=torch.float64)
        self._scalars = torch.zeros(3, device=dev, dtype=torch.float32)

    @property
    def normalization_const(self) -> torch.Tensor:
        return self.stats[6]

    def as_dict(self):
        return {n: self.stats[i] for i, n in enumerate(STAT_NAMES)}

    @torch.no_grad()
    def update(self, loss: torch.Tensor, weight_mean: torch.Tensor, weight_std: torch.Tensor, itr: int):
        torch.stack([loss.detach().reshape(()).float(), weight_mean.detach().reshape(()).float(),
                     weight_std.detach().reshape(()).float()], out=self._scalars)
        tensors = []
        for p, e in zip(self.params, self.ema_grad):
            if p.grad is None:
                raise _lib.SocmError("TrainingStatistics.update: a parameter has no gradient (call backward() first)")
            t = _lib.EmaTensor()
            t.grad, t.ema_grad, t.n = p.grad.data_ptr(), e.data_ptr(), p.numel()
            tensors.append(t)
        arr = (_lib.EmaTensor * len(tensors))(*tensors)
        _lib.check(_lib.load().socm_ema_stats_f32(arr, len(tensors), self._scalars.data_ptr(), self.stats.data_ptr(),
                                                  self._scratch.data_ptr(), int(itr), self.ema_coeff,
                                                  self.ema_weight_mean_coeff, _lib.stream_ptr()))


class Trainer:
    """One object per (solver, algorithm): ``step(itr)`` is the loop body of main.py:279-393.

    ``normalization_const`` is the starting value of the running normalisation constant that main.py:316-323 divides the
    loss by; the reference initialises it with the Monte-Carlo estimate of main.py:118-121
    (``soc_matching_b200.normalization_constant``).  With the default 1.0 the bias-corrected EMA (utils.py:389-396) jumps
    to mean(w) after the first iteration and the reported loss rescales accordingly."""

    # main.py:316-322: which objectives are divided by the normalisation constant (variance: by its square)
    _NORMALISED = ("SOCM_const_M", "SOCM_exp", "SOCM", "SOCM_adjoint", "cross_entropy")

    def __init__(self, solver, optimizer: Optional[torch.optim.Optimizer], algorithm: str, batch_size: int,
                 normalization_const=1.0, ema_coeff: float = 0.01, ema_weight_mean_coeff: float = 0.002, **loss_kw):
        self.solver, self.algorithm, self.batch_size, self.loss_kw = solver, algorithm, int(batch_size), loss_kw
        self.optimizer = optimizer
        self.statistics = TrainingStatistics(solver.neural_sde.nabla_V.parameters(), normalization_const, ema_coeff,
                                             ema_weight_mean_coeff)

    def step(self, itr: int):
        """Returns (loss / normalisation, mean(w), std(w)) as device scalars; statistics in ``self.statistics``."""
        out = self.solver.loss(self.batch_size, algorithm=self.algorithm, **self.loss_kw)
        loss, weight_mean, weight_std = out[0], out[5], out[6]
        nc = self.statistics.normalization_const.detach().clone()   # value BEFORE this iteration's EMA update
        if self.algorithm in self._NORMALISED:
            loss = loss / nc
        elif self.algorithm == "variance":
            loss = loss / nc**2
        loss.backward()                                               # main.py:323
        self.statistics.update(loss, weight_mean, weight_std, itr)   # main.py:325-345, 354-393
        if self.optimizer is not None:
            if isinstance(self.optimizer, FusedAdam):
                self.optimizer.step(zero_grad=True)                   # main.py:348-349 in one launch
            else:
                self.optimizer.step()
                self.optimizer.zero_grad()
        return loss.detach(), weight_mean, weight_std
