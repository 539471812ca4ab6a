"""Two ranks on two GPUs (NCCL) give the same objective, gradients and weight statistics as one rank
on the concatenated batch: path-indexed Philox noise + one flat all-reduce.
Skipped when fewer than 2 GPUs are visible."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["SOCM_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SOCM_ROOT"], "tests"))
from helpers import make_product_sde, random_setting, seeded_mnet, seeded_unet
import soc_matching_b200 as sb
from soc_matching_b200 import dist as sdist
rank, world, local = sdist.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
d, K, B = 10, 30, 700
st = random_setting("double_well", d, seed=2)
gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
torch.manual_seed(123)                      # same Philox key on every rank
sde = make_product_sde(st, seeded_unet(d, [256, 128, 64], 3), seeded_mnet(d, [128, 128], 4), gam, [256, 128, 64], [128, 128], dev)
solver = sb.SOC_Solver(sde, torch.zeros(d, device=dev), None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sde.sigma)
val, mw, sw = sdist.sharded_loss_backward(solver, B, os.environ.get("SOCM_ALGO", "SOCM"))
if rank == 0:
    torch.save({"val": val.cpu(), "mw": mw.cpu(), "sw": sw.cpu(),
                "grads": {n: p.grad.cpu() for n, p in sde.named_parameters() if p.grad is not None}}, os.environ["SOCM_OUT"])
if world > 1:
    torch.distributed.barrier(); torch.distributed.destroy_process_group()
'''


def _run(world, out, algo="SOCM"):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, SOCM_ROOT=ROOT, SOCM_OUT=out, SOCM_ALGO=algo)
    if world == 1:
        cmd = [sys.executable, "-c", WORKER]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", WORKER]
    subprocess.run(cmd, env=env, check=True, timeout=600)
    return torch.load(out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("algo", ["SOCM", "log-variance", "moment"])
def test_two_gpu_ranks_equal_one(tmp_path, algo):
    """log-variance is a functional of the whole batch: the ranks exchange its two moments (dist.sharded_loss_backward)."""
    from helpers import rel_l2
    one = _run(1, str(tmp_path / "one.pt"), algo)
    two = _run(2, str(tmp_path / "two.pt"), algo)
    assert abs(float(two["val"]) - float(one["val"])) <= 1e-5 * abs(float(one["val"]))
    assert abs(float(two["mw"]) - float(one["mw"])) <= 1e-6 * abs(float(one["mw"]))
    for n, g in one["grads"].items():
        assert rel_l2(two["grads"][n], g) <= 2e-5, (n, rel_l2(two["grads"][n], g))
