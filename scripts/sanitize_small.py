"""Small-shape driver for compute-sanitizer: one pass through every kernel family of the path --
tcgen05 rollout / target GEMMs / K3 on both engines (fp16 split, 3xTF32), FFMA tile kernels, the grouped stopping-time target, the tabulated-control
rollout, fused Adam and the EMA statistics.
    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_product_sde, random_setting, seeded_mnet, seeded_unet
import soc_matching_b200 as sb

DEV = "cuda"
hd = [256, 128, 64]
gam = {"gamma": torch.tensor([2.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
from soc_matching_b200 import simulate
for kind, d, K, B, stopping, flags in [("double_well", 10, 40, 130, False, ("tc", "f16")), ("double_well", 10, 40, 130, False, ("tc", "tf32")),
                                       ("double_well", 10, 40, 130, False, ("ffma",)),
                                       # B >= 1024: the target GEMMs run on the tensor cores too (target_h.cu / target_tc.cu)
                                       ("double_well", 10, 12, 1100, False, ("tc", "f16")), ("double_well", 10, 12, 1100, False, ("tc", "tf32")),
                                       ("ou_quadratic", 20, 4, 70, False, ("tc",)), ("molecular_dynamics", 1, 12, 140, True, ("tc",))]:
    simulate.ENGINE = "f16" if "f16" in flags else ("tf32" if "tf32" in flags else None)
    st = random_setting(kind, d, seed=1)
    hm = [64, 64] if stopping else [128, 128]
    sde = make_product_sde(st, seeded_unet(d, hd, 1), seeded_mnet(d, hm, 2, 0.1, 3 if stopping else 2), gam, hd, hm, DEV,
                           stopping=stopping)
    x0 = (-torch.ones(d) if stopping else torch.zeros(d)).to(DEV)
    solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sde.sigma)
    solver.force_tc, solver.force_ffma = "tc" in flags, "ffma" in flags
    opt = sb.FusedAdam([{"params": list(sde.nabla_V.parameters())}, {"params": list(sde.M.sigmoid_layers.parameters()), "lr": 1e-4}],
                       lr=1e-5)
    tr = sb.Trainer(solver, opt, "SOCM", B, normalization_const=1.0, use_stopping_time=stopping)
    for itr in range(2):
        loss, wm, ws = tr.step(itr)
    torch.cuda.synchronize()
    print(kind, flags, "loss", float(loss), "mean w", float(wm))
    for algo in ("SOCM_const_M", "SOCM_adjoint"):
        out = solver.loss(B, algorithm=algo)
        out[0].backward()
    torch.cuda.synchronize()
simulate.ENGINE = None
# tabulated control
st = random_setting("ou_quadratic", 6, seed=5)
sde = make_product_sde(st, seeded_unet(6, [16, 8, 8], 1), seeded_mnet(6, [8, 8], 2), gam, [16, 8, 8], [8, 8], DEV)
sde.use_learned_control = False
sde.u = sb.LinearControl(0.3 * torch.randn(11, 6, 6, device=DEV), 1.0)
out = sb.stochastic_trajectories(sde, torch.zeros(50, 6, device=DEV), torch.linspace(0, 1, 11, device=DEV), 1.0)
torch.cuda.synchronize()
print("tabulated rollout ok", float(out[0].abs().max()))
print("sanitize_small done")
