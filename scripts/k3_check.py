"""tcgen05 K3 vs FFMA K3: per-tensor gradient differences on the same rollout (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_product_sde, random_setting, seeded_mnet, seeded_unet, rel_l2
import soc_matching_b200 as sb
DEV = "cuda"

def run(kind, d, K, B, algo="SOCM", bench=False):
    st = random_setting(kind, d, seed=d + K)
    hd, hm = [256, 128, 64], [128, 128]
    unet, mnet = seeded_unet(d, hd, 21 + d), seeded_mnet(d, hm, 22 + d, 0.1)
    gam = {"gamma": torch.tensor([2.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    x0 = torch.zeros(d) if kind == "double_well" else 0.3 * torch.ones(d)
    noises = torch.randn(K, B, d, generator=torch.Generator().manual_seed(5))
    res = {}
    for name in ("ffma", "tc"):
        sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
        solver = sb.SOC_Solver(sde, x0.to(DEV), None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sde.sigma)
        solver.force_ffma = name == "ffma"
        solver.inject_noise(noises.to(DEV))
        out = solver.loss(B, algorithm=algo)
        out[0].backward()
        res[name] = (float(out[0]), {n: p.grad.clone() for n, p in sde.nabla_V.named_parameters()},
                     {n: p.grad.clone() for n, p in sde.M.sigmoid_layers.named_parameters()})
        if bench:
            solver.kernel_events = {}
            for _ in range(2):
                solver.loss(B, algorithm=algo)
            torch.cuda.synchronize()
            print(name, {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 3) for k, v in solver.kernel_events.items()})
    print(f"{kind} d={d} K={K} B={B}: loss ffma {res['ffma'][0]:.8g} tc {res['tc'][0]:.8g}")
    for n in res["ffma"][1]:
        print(f"   {n:18s} {rel_l2(res['tc'][1][n], res['ffma'][1][n]):.2e}")
    for n in res["ffma"][2]:
        print(f"   M {n:16s} {rel_l2(res['tc'][2][n], res['ffma'][2][n]):.2e}")

if len(sys.argv) > 1 and sys.argv[1] == "bench":
    run("double_well", 10, 200, 8192, bench=True)
else:
    run("double_well", 10, 60, 70)
    run("double_well", 10, 60, 128)
    run("ou_quadratic", 20, 12, 40)
