"""The block upper-triangular table L of the SOCM target.

The reference fills two (K+1, K+1, d, d) tensors with M(t_i, s_j) and d/ds M(t_i, s_j) by a
Python loop over rows (method.py:517-582) and later contracts them against per-sample
vectors through a (K+1, K+1, B, d, d) intermediate (method.py:591-631).  Here the same numbers
are laid out once as the left operand of a GEMM:

    L[(i,k), :] = [ M_i0[k,:]  dM_i0[k,:]  M_i1[k,:]  dM_i1[k,:] ... M_i,K-1  dM_i,K-1  M_iK[k,:] ]

with zero blocks for j < i, so that  target = R L^T  (csrc/target.cu).  Everything in this file
is torch ops (device-agnostic, differentiable w.r.t. the M-network and gamma); it is
B-independent and runs once per iteration.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class PairGrid:
    """(t, s) pairs with s >= t and the gather map from blocks of L to rows of the pair list."""

    K: int
    t: torch.Tensor        # (P,)
    s: torch.Tensor        # (P,)
    block_index: torch.Tensor  # (K+1, 2K+1) int64 into cat([M_all, dM_all, zero]) rows
    P: int
    inverse_index: torch.Tensor = None  # (2P+1,) position of every source row in the block grid


def make_pair_grid(ts: torch.Tensor, T: float) -> PairGrid:
    """s values are regenerated per row with linspace exactly like method.py:535-547 (they differ
    from ts[i:] in the last ulp at some entries), on the CPU in fp32, then moved to ts.device."""
    K = ts.shape[0] - 1
    ts_cpu = ts.detach().float().cpu()
    t_rows, s_rows = [], []
    for k in range(K + 1):
        s_rows.append(torch.linspace(ts_cpu[k], T, K + 1 - k))
        t_rows.append(ts_cpu[k] * torch.ones(K + 1 - k))
    t_vec, s_vec = torch.cat(t_rows), torch.cat(s_rows)
    P = t_vec.shape[0]
    # row offset of pair (i, j>=i) in the concatenated list
    start = torch.zeros(K + 1, dtype=torch.int64)
    for k in range(1, K + 1):
        start[k] = start[k - 1] + (K + 2 - k)
    zero_row = 2 * P
    idx = torch.full((K + 1, 2 * K + 1), zero_row, dtype=torch.int64)
    for i in range(K + 1):
        j = torch.arange(i, K + 1)
        rows = start[i] + (j - i)
        jm = j[j < K]
        idx[i, 2 * jm] = rows[: jm.shape[0]]            # M_ij
        idx[i, 2 * jm + 1] = P + rows[: jm.shape[0]]    # dM_ij
        idx[i, 2 * K] = rows[-1]                        # M_iK
    n_blocks = (K + 1) * (2 * K + 1)
    inv = torch.full((2 * P + 1,), n_blocks, dtype=torch.int64)    # default: the appended zero slot
    flat = idx.reshape(-1)
    used = flat != zero_row
    inv[flat[used]] = torch.arange(n_blocks)[used]
    dev = ts.device
    return PairGrid(K, t_vec.to(dev), s_vec.to(dev), idx.to(dev), P, inv.to(dev))


class _GatherBlocks(torch.autograd.Function):
    """blocks = cat([M, dM, 0])[block_index]; every source row lands in at most one block, so the
    backward is a gather through the inverse map (no sort-based index_put accumulation)."""

    @staticmethod
    def forward(ctx, src, block_index, inverse_index):
        ctx.save_for_backward(inverse_index)
        return src[block_index]

    @staticmethod
    def backward(ctx, gblocks):
        (inverse_index,) = ctx.saved_tensors
        flat = gblocks.reshape(-1, *gblocks.shape[2:])
        flat = torch.cat([flat, flat.new_zeros(1, *flat.shape[1:])], dim=0)   # slot for unused rows
        return flat[inverse_index], None, None


def build_L(m_all: torch.Tensor, dm_all: torch.Tensor, grid: PairGrid, ldr: int) -> torch.Tensor:
    """(P,d,d) x2 -> L ((K+1)d, ldr) fp32, zero-padded to the row pitch of R."""
    K, d = grid.K, m_all.shape[-1]
    zero = torch.zeros(1, d, d, device=m_all.device, dtype=m_all.dtype)
    src = torch.cat([m_all, dm_all, zero], dim=0)
    blocks = _GatherBlocks.apply(src, grid.block_index, grid.inverse_index)   # (K+1, 2K+1, d, d)
    L = blocks.permute(0, 2, 1, 3).reshape((K + 1) * d, (2 * K + 1) * d)
    if ldr > L.shape[1]:
        L = torch.nn.functional.pad(L, (0, ldr - L.shape[1]))
    return L.contiguous()


def build_LT_grouped(m_all: torch.Tensor, dm_all: torch.Tensor, grid: PairGrid, nrp: int) -> torch.Tensor:
    """Stopping-time variant: (P, Q, d, d) x2 (one table per stopping index q, method.py:524-564) ->
    LT (Q, (2K+1)d, nrp) fp32, the transposed block table of csrc/target_grouped.cu:
        LT[q][(j', l)][(i, k)] = block (i, j') of [M_i0 dM_i0 ... M_iK] for index q, entry (k, l),
    zero left of the block diagonal and in the row padding."""
    K, d, Q = grid.K, m_all.shape[-1], m_all.shape[1]
    zero = torch.zeros(1, Q, d, d, device=m_all.device, dtype=m_all.dtype)
    src = torch.cat([m_all, dm_all, zero], dim=0)                                  # (2P+1, Q, d, d)
    blocks = _GatherBlocks.apply(src, grid.block_index, grid.inverse_index)        # (K+1, 2K+1, Q, d(k), d(l))
    LT = blocks.permute(2, 1, 4, 0, 3).reshape(Q, (2 * K + 1) * d, (K + 1) * d)
    if nrp > LT.shape[2]:
        LT = torch.nn.functional.pad(LT, (0, nrp - LT.shape[2]))
    return LT.contiguous()
