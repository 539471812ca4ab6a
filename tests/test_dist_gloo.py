"""Data-parallel plumbing (soc_matching_b200/dist.py) on CPU: world_size 2, gloo backend.
The CUDA kernels cannot run here, so the per-rank ``solver.loss`` is a torch stand-in whose value and
gradients are a deterministic function of the GLOBAL path index (like the path-indexed Philox draws of
the real rollout); the test checks what the plumbing must guarantee: sharding bounds, one flat
all-reduce, and that the 2-rank objective / gradients / weight statistics equal the 1-rank run."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


class _FakeSDE(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Parameter(torch.randn(7))
        self.b = torch.nn.Parameter(torch.randn(3, 5))


class _FakeSolver:
    """Mimics the parts of SOC_Solver that dist.sharded_loss_backward touches."""

    def __init__(self):
        self.neural_sde = _FakeSDE()
        self.path_offset = 0
        self.last_stats = None
        self.batch_reduce = None
        self.global_batch = None

    def parameters(self):
        return self.neural_sde.parameters()

    def loss(self, batch, algorithm="SOCM", use_stopping_time=False):
        idx = torch.arange(self.path_offset, self.path_offset + batch, dtype=torch.float64)
        w = torch.exp(-0.001 * idx)                                    # per-path importance weight
        feat = torch.stack([torch.sin(idx * (k + 1) * 0.01) for k in range(7)], 1).float()
        per_path = (feat @ self.neural_sde.a) ** 2 + (self.neural_sde.b.sum() * torch.cos(idx * 0.02).float()) ** 2
        if algorithm == "log-variance":                                # a variance over ALL paths (solver.py, method.py:800-856)
            s_m = feat @ self.neural_sde.a + self.neural_sde.b.sum() * torch.cos(idx * 0.02).float()
            m1, m2 = s_m.sum(), (s_m * s_m).sum()
            n = batch
            if self.batch_reduce is not None:
                tot = self.batch_reduce(torch.stack([m1.detach(), m2.detach()]).double())
                m1 = m1 + (tot[0] - m1.detach().double()).float()
                m2 = m2 + (tot[1] - m2.detach().double()).float()
                n = self.global_batch
            obj = n / (n - 1) * (m2 / n - (m1 / n) ** 2)
            self.last_stats = torch.stack([w.sum(), (w * w).sum(), torch.tensor(float(batch), dtype=torch.float64)])
            return (obj, None, None, None, None, None, None, None)
        alive = 1.0 + (idx % 5)                                        # "sum of stop indicators" of each path
        z = alive.sum() if use_stopping_time else torch.tensor(float(batch), dtype=torch.float64)
        obj = (per_path * w.float()).sum() / z.float()                 # shard-normalised, like method.py:715 / 720
        self.last_stats = torch.stack([w.sum(), (w * w).sum(), z])
        return (obj, None, None, None, None, None, None, None)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, out, stopping=False, algorithm="SOCM"):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from soc_matching_b200 import dist as sdist
    r, w, _ = sdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    solver = _FakeSolver()
    kw = {"use_stopping_time": True} if stopping else {}
    val, mean_w, std_w = sdist.sharded_loss_backward(solver, batch, algorithm, **kw)
    if rank == 0:
        out.put((float(val), float(mean_w), float(std_w), solver.neural_sde.a.grad.clone(), solver.neural_sde.b.grad.clone()))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    from soc_matching_b200.dist import shard_bounds
    for n in (0, 1, 7, 128, 1 << 20, 1000003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
@pytest.mark.parametrize("stopping,algorithm", [(False, "SOCM"), (True, "SOCM"), (False, "log-variance")])
def test_two_ranks_equal_one_rank(stopping, algorithm):
    """stopping=True: the objective is normalised by the GLOBAL sum of stop indicators (method.py:715), so the shards
    are weighted by z_shard / z_total, not by their path counts.  log-variance: a variance over all paths -- the two
    moments are summed over the ranks and the gradients of the ranks add up without shard weights."""
    batch = 1001                                                        # ragged: 501 + 500 paths
    from soc_matching_b200 import dist as sdist
    ref = _FakeSolver()
    kw = {"use_stopping_time": True} if stopping else {}
    val1, mean1, std1 = sdist.sharded_loss_backward(ref, batch, algorithm, **kw)  # no process group: world = 1
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, out, stopping, algorithm)) for r in range(2)]
    for p in procs:
        p.start()
    val2, mean2, std2, ga, gb = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(val2 - float(val1)) <= 1e-5 * abs(float(val1))
    assert abs(mean2 - float(mean1)) <= 1e-6 and abs(std2 - float(std1)) <= 1e-6
    assert torch.allclose(ga, ref.neural_sde.a.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(gb, ref.neural_sde.b.grad, rtol=1e-4, atol=1e-6)


def test_more_ranks_than_paths_is_rejected():
    from soc_matching_b200 import dist as sdist
    solver = _FakeSolver()
    sdist.sharded_loss_backward(solver, 1, "SOCM")      # world = 1: fine
    assert all(p.grad is not None for p in solver.parameters())
