"""Is the error of the tcgen05 UNet forward a systematic shrink (truncating accumulation)?  One Euler step (K=1)
from real mid-rollout states: controls = -nabla_V(t0, x) from the tcgen05 and the FFMA kernels vs torch fp64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from helpers import make_product_sde, random_setting, seeded_mnet, seeded_unet
from soc_matching_b200 import simulate
DEV = "cuda"
d, K, B = 10, 200, 75776
st_ = random_setting("double_well", d, seed=4)
hd, hm = [256, 128, 64], [128, 128]
gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
sde = make_product_sde(st_, seeded_unet(d, hd, 5), seeded_mnet(d, hm, 6, 0.1), gam, hd, hm, DEV)
ts = torch.linspace(0, 1, K + 1, device=DEV)
wsp = simulate.rollout(sde, torch.zeros(B, d, device=DEV), ts, 1.0, seed=99)
x = wsp.states[120].clone()
t2 = ts[120:122].clone()
noise = torch.zeros(1, B, d, device=DEV)
u_tc = simulate.rollout(sde, x, t2, 1.0, noises=noise).controls[0].double()
u_ff = simulate.rollout(sde, x, t2, 1.0, noises=noise, force_ffma=True).controls[0].double()
unet = sde.nabla_V
names = ["down_0", "down_1", "down_2", "res_0", "res_1", "res_2", "up_2", "up_1", "up_0"]
P = {n + s: getattr(getattr(unet, n)[0], s[1:]).detach().double() for n in names for s in (".weight", ".bias")}
lin = lambda n, v: F.linear(v, P[n + ".weight"], P[n + ".bias"])
tx = torch.cat([t2[0].double().reshape(1, 1).expand(B, 1), x.double()], -1)
r1 = torch.relu(lin("down_0", tx)); r2 = torch.relu(lin("down_1", r1)); r3 = torch.relu(lin("down_2", r2))
o2 = torch.relu(lin("up_2", r3)) + lin("res_2", r2)
o1 = torch.relu(lin("up_1", o2)) + lin("res_1", r1)
u64 = -(torch.relu(lin("up_0", o1)) + lin("res_0", tx))
for name, u in (("tc", u_tc), ("ffma", u_ff)):
    e = u - u64
    slope = float((e * u64).sum() / (u64 * u64).sum())
    resid = e - slope * u64
    print(f"{name}: rel err {float(e.norm() / u64.norm()):.2e}  slope on truth {slope:+.2e}  "
          f"rel err after removing the slope {float(resid.norm() / u64.norm()):.2e}")
