"""Data parallelism over trajectories: one process per GPU, rank r rolls out
paths [r B/W, (r+1) B/W) with path-indexed Philox counters (results do not depend on W), and a
single all-reduce per iteration carries the parameter gradients and the scalar normalisers.
The reference has no distributed code at all (one GPU per job, configs/soc.yaml:53-62)."""
from __future__ import annotations

import math
import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the process group
    when WORLD_SIZE > 1 (NCCL on GPUs, gloo on CPU)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_bounds(n_paths: int, rank: int, world: int) -> tuple:
    """[lo, hi) of the paths owned by ``rank``; shards differ by at most one path."""
    base, rem = divmod(n_paths, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat(tensors: Iterable[torch.Tensor], group=None) -> List[torch.Tensor]:
    """Sum-all-reduce a list of tensors through ONE flat buffer (one collective, latency-bound at the
    ~0.8 MB this path moves) and write the results back in place."""
    tensors = list(tensors)
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].reshape(t.shape).to(t.dtype))
        off += n
    return tensors


def merge_weight_stats(stats: torch.Tensor, n_paths: int):
    """(mean, unbiased std) of the importance weights from the (all-reduced) fp64 sums
    [sum w, sum w^2, ...]  (method.py:903-904)."""
    mean = stats[0] / n_paths
    var = (stats[1] - stats[0] * stats[0] / n_paths) / max(n_paths - 1, 1)
    return mean.float(), torch.sqrt(torch.clamp(var, min=0.0)).float()


def sharded_loss_backward(solver, global_batch: int, algorithm: str = "SOCM", group=None, **loss_kw):
    """One data-parallel SOCM iteration: every rank runs ``solver.loss`` on its shard and
    back-propagates; gradients of all parameters and the normalisers are summed with one
    all-reduce.  Returns (global objective, mean(w), std(w)).

    Without stopping times the objective of a shard is normalised by its own size (K+1) n_shard, so the
    global value is the shard-size weighted mean.  With ``use_stopping_time`` (SOCM only, method.py:711-715)
    the reference divides by the sum of ALL stop indicators: each shard is normalised by its own sum z_r, so
    the ranks first all-reduce z (one fp64 scalar) and weight their shard by z_r / z before back-propagating --
    the result does not depend on the world size.

    ``log-variance`` and ``variance`` are functionals of the whole batch (a variance over all paths, method.py:800-856):
    the solver sums the two moments over the ranks (``solver.batch_reduce``) and every rank differentiates the same global
    value with respect to its own paths, so the gradients add up (no shard weighting) and the value is not summed."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > global_batch:
        raise ValueError(f"world size {world} exceeds the global batch {global_batch}: a rank would own no path")
    lo, hi = shard_bounds(global_batch, rank, world)
    solver.path_offset = lo
    whole_batch = algorithm in ("log-variance", "variance")

    def _sum_over_ranks(t):
        if world > 1:
            dist.all_reduce(t, group=group)
        return t

    if whole_batch:
        solver.batch_reduce, solver.global_batch = _sum_over_ranks, global_batch
    try:
        out = solver.loss(hi - lo, algorithm=algorithm, **loss_kw)
    finally:
        if whole_batch:
            solver.batch_reduce, solver.global_batch = None, None
    stats = solver.last_stats.clone()
    share = (hi - lo) / float(global_batch)
    if whole_batch:
        share = 1.0
    if loss_kw.get("use_stopping_time") and algorithm == "SOCM":
        z_all = stats[2:3].clone()
        if world > 1:
            dist.all_reduce(z_all, group=group)
        share = float(stats[2] / z_all[0])
    (out[0] * share).backward()
    # Deterministic reduction list: every parameter of the solver (the neural SDE's networks and gammas plus the
    # solver's own y0 / gamma of SOCM_exp, main.py:166), in registration order, with a zero standing in for a
    # gradient this rank did not produce -- all ranks then send flat buffers of the same length.
    params = list(solver.parameters())
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    value = (out[0].detach() * (1.0 / world if whole_batch else share)).reshape(1)
    if world > 1:
        allreduce_flat(grads + [value], group)
        dist.all_reduce(stats, group=group)  # fp64 sums stay fp64
        for p, g in zip(params, grads):
            if p.grad is None and bool(torch.any(g != 0)):
                p.grad = g
    mean_w, std_w = merge_weight_stats(stats, global_batch)
    return value[0], mean_w, std_w
