"""Which kernel carries the tcgen05-vs-fp32 difference of a full-chunk SOCM iteration?  (1) all tcgen05,
(2) all FFMA/SIMT, (3) tcgen05 rollout + FFMA/SIMT target and loss kernels, same Philox key."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_product_sde, random_setting, rel_l2, seeded_mnet, seeded_unet
import soc_matching_b200 as sb
from soc_matching_b200 import simulate
DEV = "cuda"
d, K, B = 10, 200, int(os.environ.get("AB_B", 75776))
st = random_setting("double_well", d, seed=4)
hd, hm = [256, 128, 64], [128, 128]
unet, mnet = seeded_unet(d, hd, 5), seeded_mnet(d, hm, 6, 0.1)
gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
orig_rollout = simulate.rollout
def run(ffma_loss, ffma_rollout, simt_target=False):
    torch.manual_seed(1234)
    simulate._SEED_COUNTER[0] = 77
    sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV)
    solver = sb.SOC_Solver(sde, torch.zeros(d, device=DEV), None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sde.sigma)
    solver.force_ffma = ffma_loss
    solver.force_simt_target = simt_target
    def patched(*a, **kw):
        kw["force_ffma"] = ffma_rollout
        return orig_rollout(*a, **kw)
    simulate.rollout = patched
    out = solver.loss(B, algorithm="SOCM")
    out[0].backward()
    simulate.rollout = orig_rollout
    g = {n: q.grad.clone() for n, q in sde.named_parameters() if q.grad is not None}
    return float(out[0].detach()), g
r_tc = run(False, False)
r_ff = run(True, True)
r_mix = run(True, False)
def cmp(a, b, tag):
    print(tag, "loss rel %.2e" % (abs(a[0] - b[0]) / abs(b[0])))
    worst = sorted(((rel_l2(a[1][n], b[1][n]), n) for n in b[1]), reverse=True)[:4]
    print("   worst tensors:", [(f"{v:.2e}", n) for v, n in worst])
r_k2 = run(False, False, simt_target=True)
cmp(r_tc, r_k2, "tc K2 vs SIMT K2 (tcgen05 K1, K3 in both):          ")
cmp(r_k2, r_mix, "tc K3 vs FFMA K3 (tcgen05 K1, SIMT K2 in both):     ")
cmp(r_tc, r_mix, "tc K2/K3 vs fp32 K2/K3 on the SAME tcgen05 rollout:")
cmp(r_mix, r_ff, "tcgen05 rollout vs FFMA rollout, same fp32 K2/K3:   ")
cmp(r_tc, r_ff, "all tcgen05 vs all fp32:                           ")
