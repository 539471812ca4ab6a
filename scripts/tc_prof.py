"""Short tcgen05 rollout run for ncu: double_well d=10, K=50, B = 2 tiles per SM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_product_sde, random_setting, seeded_mnet, seeded_unet
from soc_matching_b200 import simulate
DEV = "cuda"
d, K, B = 10, int(os.environ.get("K", 50)), 148 * 128 * int(os.environ.get("TILES", 2))
st = random_setting("double_well", d, seed=3)
gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
sde = make_product_sde(st, seeded_unet(d, [256, 128, 64], 5), seeded_mnet(d, [128, 128], 6), gam, [256, 128, 64], [128, 128], DEV)
ts = torch.linspace(0, 1.0, K + 1, device=DEV)
x0 = torch.zeros(B, d, device=DEV)
for _ in range(2):
    simulate.rollout(sde, x0, ts, 1.0, seed=1)
torch.cuda.synchronize()
import ctypes
from soc_matching_b200 import _lib
lib = _lib.load()
if hasattr(lib, "socm_debug_tc_prof"):
    buf = (ctypes.c_ulonglong * 48)()
    lib.socm_debug_tc_prof(buf)
    steps = K * int(os.environ.get("TILES", 2))
    names_e = ["xin", "noise", "wait D0", "r1 chunks (down_1 + Wc)", "wait D1", "epi(r2,r3,y2,o2 work)", "wait D2/D3A", "wait D3B", "wait D4A", "-", "-", "y1 chunks (folded up_0)", "wait Y0", "y0 read", "sde", "exchange sync"]
    names_m = ["loop", "wait XIN", "down0+down_1(+Wc) issue", "wait R2", "issue(d2,u2,r2)", "wait R3/Y2", "wait O2", "issue up_1", "-", "-", "-", "issue up_0 (y1 chunks)"]
    print("E thread 0 (cycles per step):")
    for i, n in enumerate(names_e):
        print(f"  {n:28s} {buf[i] / steps:9.0f}")
    print("  total", sum(buf[:16]) / steps)
    print("helper thread 128 (cycles per step):")
    for i, n in enumerate(names_e):
        print(f"  {n:28s} {buf[32 + i] / steps:9.0f}")
    print("M warp (cycles per step):")
    for i, n in enumerate(names_m):
        print(f"  {n:28s} {buf[16 + i] / steps:9.0f}")
    print("  total", sum(buf[16:]) / steps)
