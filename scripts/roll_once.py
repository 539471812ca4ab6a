"""One rollout call for ncu captures: ENGINE=f16|tf32 B=75776 python scripts/roll_once.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import soc_matching_b200 as sb
from soc_matching_b200 import simulate
B, K, d = int(os.environ.get("B", 75776)), int(os.environ.get("K", 200)), 10
simulate.ENGINE = os.environ.get("ENGINE", "f16")
torch.manual_seed(0)
x0, sigma, sde = sb.make_benchmark_sde("double_well", d, device="cuda", gamma=6.0, scaling_factor_M=0.1)
ts = torch.linspace(0, 1, K + 1, device="cuda")
for _ in range(2):
    w = simulate.rollout(sde, x0.repeat(B, 1), ts, 1.0, seed=5)
torch.cuda.synchronize()
print("ok", float(w.lw.sum()))
