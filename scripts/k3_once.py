"""One SOCM iteration at one chunk for ncu captures: ENGINE=f16|tf32 B=75776 python scripts/k3_once.py"""
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
solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sigma)
for _ in range(int(os.environ.get("ITERS", 1))):
    out = solver.loss(B, algorithm="SOCM")
    out[0].backward()
torch.cuda.synchronize()
print("ok", float(out[0]))
