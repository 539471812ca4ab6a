"""Soak run: N SOCM training iterations (Trainer: loss -> backward -> FusedAdam -> EMA statistics) at one full chunk on the
default dispatch; checks that everything stays finite and prints the loss trace.  python scripts/soak.py [iters] [B]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import soc_matching_b200 as sb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = int(sys.argv[2]) if len(sys.argv) > 2 else 75776
torch.manual_seed(0)
d, K = 10, 200
x0, sigma, sde = sb.make_benchmark_sde("double_well", d, device="cuda", gamma=6.0, scaling_factor_M=0.1)
solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sigma)
mode = os.environ.get("SOAK_MODE", "default")   # default | ffma | tf32
solver.force_ffma = mode == "ffma"
if mode == "tf32":
    from soc_matching_b200 import simulate
    simulate.ENGINE = "tf32"
lr_m = float(os.environ.get("SOAK_LR_M", 1e-3))
opt = sb.FusedAdam([{"params": list(sde.nabla_V.parameters())},
                    {"params": list(sde.M.sigmoid_layers.parameters()), "lr": lr_m}, {"params": [sde.gamma], "lr": lr_m}], lr=1e-4)
tr = sb.Trainer(solver, opt, "SOCM", B, normalization_const=1.0)
t0 = time.time()
trace = []
for itr in range(iters):
    loss, wm, ws = tr.step(itr)
    trace.append(float(loss))
    assert all(torch.isfinite(p).all() for p in solver.parameters()), f"non-finite parameter at iteration {itr}"
torch.cuda.synchronize()
print(mode, "trace", " ".join(f"{x:.1f}" for x in trace[::max(1, iters // 12)]))
print(f"{iters} iterations at B={B}: {time.time() - t0:.1f} s; loss first/last {trace[0]:.4f} / {trace[-1]:.4f}; "
      f"min {min(trace):.4f} max {max(trace):.4f}; mean w {float(wm):.4f}")
assert all(x == x and abs(x) < 1e30 for x in trace)
print("soak ok")
