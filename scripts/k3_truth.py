"""K3 (tcgen05 and fp32 FFMA) against torch fp64 autograd on REAL rollout states at bench-chunk size:
per-tensor relative L2 error of each kernel's UNet gradient."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from helpers import make_product_sde, random_setting, seeded_mnet, seeded_unet
from soc_matching_b200 import _lib, networks, simulate
DEV = "cuda"
torch.manual_seed(0)
d, K = 10, 200
B = int(os.environ.get("AB_B", 75776))
lib = _lib.load()
st_ = random_setting("double_well", d, seed=4)
hd, hm = [256, 128, 64], [128, 128]
unet_p, mnet_p = seeded_unet(d, hd, 5), seeded_mnet(d, hm, 6, 0.1)
gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
sde = make_product_sde(st_, unet_p, mnet_p, gam, hd, hm, DEV)
ts = torch.linspace(0, 1, K + 1, device=DEV)
wsp = simulate.rollout(sde, torch.zeros(B, d, device=DEV), ts, 1.0, seed=99)
states = wsp.states
unet = sde.nabla_V
udesc, keep = networks.unet_desc(unet)
g = torch.Generator(DEV).manual_seed(1)
ldt = ((K + 1) * d + 3) // 4 * 4
target = 3.0 * torch.randn(B, ldt, device=DEV, generator=g)
w = torch.exp(wsp.lw[0] + wsp.lw[1] + wsp.lw[2])
if os.environ.get("REAL_TARGET", "0") == "1":      # the SOCM target of a real iteration (cancelling residual)
    import soc_matching_b200 as sb
    solver = sb.SOC_Solver(sde, torch.zeros(d, device=DEV), None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sde.sigma)
    solver._debug_keep = True
    solver.loss(B, algorithm="SOCM")
    states, target, w = solver._debug_last["states"], solver._debug_last["target"], solver._debug_last["w"]
    del solver
    torch.cuda.empty_cache()
st = _lib.Setting()
eye, kap = torch.eye(d, device=DEV), torch.ones(d, device=DEV)
st.kind, st.d, st.sigma_is_identity, st.lmbd = 2, d, 1, 1.0
st.sigma, st.sigma_inv, st.kappa, st.nu = eye.data_ptr(), eye.data_ptr(), kap.data_ptr(), kap.data_ptr()
ws = torch.zeros(int(lib.socm_loss_workspace_bytes(udesc, B, K)) // 4 + 1024, device=DEV)
scale = 1.0 / ((K + 1) * B)
out = {}
for name, flag in (("ffma", _lib.LOSS_FORCE_FFMA), ("tc", _lib.LOSS_FORCE_TC)):
    G = torch.zeros(B, ldt, device=DEV)
    grad = torch.zeros(int(lib.socm_unet_param_count(udesc)), device=DEV)
    loss = torch.zeros(1, device=DEV, dtype=torch.float64)
    _lib.check(lib.socm_unet_loss_fwdbwd_f32(st, udesc, None, ts.data_ptr(), states.data_ptr(), target.data_ptr(), ldt,
                                             w.data_ptr(), None, scale, B, K, G.data_ptr(), grad.data_ptr(),
                                             loss.data_ptr(), ws.data_ptr(), flag, _lib.stream_ptr()))
    torch.cuda.synchronize()
    out[name] = (float(loss), grad.clone())
# ---- fp64 autograd over all points, in slabs of grid times
names = ["down_0", "down_1", "down_2", "res_0", "res_1", "res_2", "up_2", "up_1", "up_0"]
P = {n + s: getattr(unet, n)[0].__getattr__(s.strip(".")).detach().double().requires_grad_(True)
     for n in names for s in (".weight", ".bias")}
lin = lambda n, v: F.linear(v, P[n + ".weight"], P[n + ".bias"])
tot = 0.0
near = 0
for i0 in range(0, K + 1, 8):
    i1 = min(K + 1, i0 + 8)
    x = states[i0:i1].double()
    tx = torch.cat([ts[i0:i1].double().reshape(-1, 1, 1).expand(i1 - i0, B, 1), x], -1)
    z1 = lin("down_0", tx); r1 = torch.relu(z1)
    z2 = lin("down_1", r1); r2 = torch.relu(z2)
    z3 = lin("down_2", r2); r3 = torch.relu(z3)
    y2 = lin("up_2", r3); o2 = torch.relu(y2) + lin("res_2", r2)
    y1 = lin("up_1", o2); o1 = torch.relu(y1) + lin("res_1", r1)
    y0 = lin("up_0", o1)
    outv = torch.relu(y0) + lin("res_0", tx)
    for z in (z1, z2, z3, y2, y1, y0):
        near += int((z.abs() < 1e-5).sum())
    tgt = target[:, i0 * d:i1 * d].reshape(B, i1 - i0, d).permute(1, 0, 2).double()
    L = (((outv - tgt) ** 2).sum(-1) * w.double()[None]).sum() * scale
    L.backward()
    tot += float(L)
print("pre-activations within 1e-5 of a kink:", near, "of", (K + 1) * B * (256 + 128 + 64 + 128 + 256 + d))
print("loss rel: ffma %.2e  tc %.2e" % (abs(out["ffma"][0] - tot) / tot, abs(out["tc"][0] - tot) / tot))
off = 0
for n in names:
    for s in (".weight", ".bias"):
        t = P[n + s].grad.flatten()
        a, b = out["tc"][1][off:off + t.numel()].double(), out["ffma"][1][off:off + t.numel()].double()
        print(f"  {n}{s}: tc {float((a - t).norm() / t.norm()):.2e}  ffma {float((b - t).norm() / t.norm()):.2e}")
        off += t.numel()
