"""Training-loop cross-check: the same SOCM iterations with (a) Trainer + FusedAdam, (b) a hand-written loop with
torch.optim.Adam; prints the loss traces (same Philox seeds).  The Trainer reports loss / normalization_const as
main.py:316-323 does (the constant is the bias-corrected EMA of mean(w), main.py:354-359), so from iteration 1 on its
trace is the plain loss divided by ~mean(w); Adam being scale-invariant, the parameters follow the same path, which the
last line checks.  python scripts/train_check.py [iters] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import soc_matching_b200 as sb
from soc_matching_b200 import simulate

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d, K = 10, 200


def make():
    torch.manual_seed(0)
    simulate._SEED_COUNTER[0] = 777
    x0, sigma, sde = sb.make_benchmark_sde("double_well", d, device="cuda", gamma=6.0, scaling_factor_M=0.1)
    solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sigma)
    groups = [{"params": list(sde.nabla_V.parameters())}, {"params": list(sde.M.sigmoid_layers.parameters()), "lr": 1e-3},
              {"params": [sde.gamma], "lr": 1e-3}]
    return sde, solver, groups


sde, solver, groups = make()
tr = sb.Trainer(solver, sb.FusedAdam(groups, lr=1e-4), "SOCM", B, normalization_const=1.0)
a = [float(tr.step(i)[0]) for i in range(iters)]
sde, solver, groups = make()
opt = torch.optim.Adam(groups, lr=1e-4)
b = []
for i in range(iters):
    opt.zero_grad()
    out = solver.loss(B, algorithm="SOCM")
    out[0].backward()
    opt.step()
    b.append(float(out[0]))
    if i == 0:
        gn = {n: float(p.grad.norm()) for n, p in sde.named_parameters() if p.grad is not None}
        print("grad norms at iteration 0:", {k: round(v, 4) for k, v in list(gn.items())[:6]}, "... max", max(gn.values()))
pa = torch.cat([p.detach().flatten() for p in tr.solver.parameters()])
pb = torch.cat([p.detach().flatten() for p in solver.parameters()])
print("Trainer + FusedAdam:", " ".join(f"{x:.2f}" for x in a))
print("loop + torch Adam  :", " ".join(f"{x:.2f}" for x in b))
print("ratio               :", " ".join(f"{x / y:.2f}" for x, y in zip(a, b)))
print("relative distance of the two parameter vectors after the run:", float((pa - pb).norm() / pb.norm()))
