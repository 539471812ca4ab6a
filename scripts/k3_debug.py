"""Dump the K3a scratch and compare each operand tensor with a torch fp64 autograd evaluation."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from helpers import seeded_unet, rel_l2
from soc_matching_b200 import _lib, networks
import soc_matching_b200 as sb
DEV = "cuda"
d, K, B = 10, int(sys.argv[2]) if len(sys.argv) > 2 else 3, int(sys.argv[1]) if len(sys.argv) > 1 else 70
lib = _lib.load()
p = {k: v.to(DEV) for k, v in seeded_unet(d, [256, 128, 64], 31).items()}
unet = networks.FullyConnectedUNet(d, (256, 128, 64), 1.0).to(DEV)
unet.load_state_dict(p)
udesc, keep = networks.unet_desc(unet)
g = torch.Generator(DEV).manual_seed(1)
states = torch.randn(K + 1, B, d, device=DEV, generator=g)
ts = torch.linspace(0, 1, K + 1, device=DEV)
ldt = ((K + 1) * d + 3) // 4 * 4
target = torch.randn(B, ldt, device=DEV, generator=g)
w = torch.exp(0.3 * torch.randn(B, device=DEV, generator=g))
G = torch.zeros(B, ldt, device=DEV)
npar = int(lib.socm_unet_param_count(udesc))
grad = torch.zeros(npar, device=DEV)
loss = torch.zeros(1, device=DEV, dtype=torch.float64)
wsb = int(lib.socm_loss_workspace_bytes(udesc, B, K))
ws = torch.zeros(wsb // 4 + 1024, device=DEV)
st = _lib.Setting()
eye = torch.eye(d, device=DEV); kap = torch.ones(d, device=DEV)
st.kind, st.d, st.sigma_is_identity, st.lmbd = 2, d, 1, 1.0
st.sigma, st.sigma_inv, st.kappa, st.nu = eye.data_ptr(), eye.data_ptr(), kap.data_ptr(), kap.data_ptr()
scale = 1.0 / ((K + 1) * B)
_lib.check(lib.socm_unet_loss_fwdbwd_f32(st, udesc, None, ts.data_ptr(), states.data_ptr(), target.data_ptr(), ldt,
                                         w.data_ptr(), None, scale, B, K, G.data_ptr(), grad.data_ptr(),
                                         loss.data_ptr(), ws.data_ptr(), 0, _lib.stream_ptr()))
torch.cuda.synchronize()
# ---- torch reference with intermediates
P = {k: v.double().requires_grad_(True) for k, v in p.items()}
tx = torch.cat([ts.reshape(-1, 1, 1).expand(K + 1, B, 1), states], -1).double()
lin = lambda n, v: F.linear(v, P[n + ".0.weight"], P[n + ".0.bias"])
inter = {}
def keep_(name, t):
    t.retain_grad(); inter[name] = t; return t
z1 = keep_("z1", lin("down_0", tx)); r1 = torch.relu(z1)
z2 = keep_("z2", lin("down_1", r1)); r2 = torch.relu(z2)
z3 = keep_("z3", lin("down_2", r2)); r3 = torch.relu(z3)
y2 = keep_("y2", lin("up_2", r3)); o2 = keep_("o2", torch.relu(y2) + lin("res_2", r2))
y1 = keep_("y1", lin("up_1", o2)); o1 = keep_("o1", torch.relu(y1) + lin("res_1", r1))
y0 = keep_("y0", lin("up_0", o1)); o0 = keep_("o0", torch.relu(y0) + lin("res_0", tx))
tgt = target[:, :(K + 1) * d].reshape(B, K + 1, d).permute(1, 0, 2).double()
L = (((o0 - tgt) ** 2).sum(-1) * w.double()[None]).sum() * scale
L.backward()
print("loss", float(loss), float(L))
ref = {"R1": r1, "R2": r2, "R3": r3, "O2": o2, "O1": o1, "DY0": y0.grad, "DO0": o0.grad, "DY1": y1.grad, "DO1": o1.grad,
       "DY2": y2.grad, "DO2": o2.grad, "DZ3": z3.grad, "DZ2": z2.grad, "DZ1": z1.grad}
FB = dict(XIN=0, R1=1, R2=9, R3=13, O2=15, O1=19, DY0=27, DO0=28, DY1=29, DO1=37, DY2=45, DO2=49, DZ3=53, DZ2=55, DZ1=59)
NFB = 67
# locate scratch
kin = ((d + 1 + 7) // 8) * 8
small_total = (1216 + 256 * kin + kin + kin * kin + kin + 3) // 4 * 4
S = 2 if kin > 16 else 1
tape = ((2 * (40 + S) * 32768 + small_total * 4) + 1023) // 1024 * 1024
base = ws.data_ptr() + tape
off = ((1024 - base % 1024) % 1024 + tape) // 4
n_mblk = (B + 127) // 128
n_tiles = (K + 1) * n_mblk
words = ws[off: off + n_tiles * 4 * NFB * 1024].reshape(n_tiles, 4, NFB, 32, 32).cpu()
r = torch.arange(32)
def extract(name, width):
    out = torch.zeros(n_tiles, 4, 32, width)
    for fbi in range(width // 32):
        blk = words[:, :, FB[name] + fbi]             # tile, q, row, word
        for u in range(4):
            phys = (u ^ (r & 3))
            idx = (phys[:, None] * 8 + torch.arange(8)[None, :])
            out[:, :, :, fbi * 32 + u * 8: fbi * 32 + u * 8 + 8] = torch.gather(blk, 3, idx[None, None].expand(n_tiles, 4, 32, 8))
    return out.reshape(n_tiles, 128, width)
for name, t in ref.items():
    width = t.shape[-1] if t.shape[-1] >= 32 else 32
    got = extract(name, width)[:, :, :t.shape[-1]]      # tile, point, feat
    got = got.reshape(K + 1, n_mblk * 128, -1)[:, :B]
    e = rel_l2(got, t.detach().cpu())
    # per-quarter errors
    eq = [rel_l2(got[:, a:b], t.detach().cpu()[:, a:b]) for a, b in ((0, 32), (32, 64), (64, min(96, B)), (96, B)) if b > a]
    print(f"{name:4s} rel {e:.2e}  per lane-quarter {['%.1e' % x for x in eq]}")
# ---- final gradients vs torch
names = ["down_0", "down_1", "down_2", "res_0", "res_1", "res_2", "up_2", "up_1", "up_0"]
o = 0
gc = grad.cpu()
for n in names:
    for suffix in (".0.weight", ".0.bias"):
        t = P[n + suffix].grad.cpu()
        got = gc[o:o + t.numel()].reshape(t.shape); o += t.numel()
        print(f"grad {n + suffix:18s} {rel_l2(got, t):.2e}")
# non-live rows must be exactly zero in every dY tensor
if B % 128:
    for name in ("DY0", "DO0", "DY1", "DO1", "DY2", "DO2", "DZ3", "DZ2", "DZ1"):
        width = max(ref[name].shape[-1], 32)
        got = extract(name, width).reshape(K + 1, n_mblk * 128, -1)[:, B:]
        print(name, "non-live max abs", float(got.abs().max()), "nan", bool(torch.isnan(got).any()))
got = extract("DY1", 256).reshape(K + 1, n_mblk * 128, -1)[:, :B]
want = ref["DY1"].detach().cpu().float()
bad = (got - want).abs() > 1e-3 * want.abs().max()
print("DY1 bad entries", int(bad.sum()), "of", bad.numel())
idx = bad.nonzero()
print("tiles:", sorted(set(idx[:, 0].tolist()))[:40])
print("points:", sorted(set(idx[:, 1].tolist()))[:40])
print("feats:", sorted(set(idx[:, 2].tolist()))[:64])
for t_, p_, f_ in idx[:10].tolist():
    print(t_, p_, f_, float(got[t_, p_, f_]), float(want[t_, p_, f_]), float(ref["DO1"][t_, p_, f_]))
