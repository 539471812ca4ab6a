"""K3 tcgen05 vs fp32 FFMA on identical inputs at bench-chunk size: per-tensor relative L2 difference of the UNet
gradients (SOCM_B200_LIB selects the library)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import seeded_unet
from soc_matching_b200 import _lib, networks
DEV = "cuda"
d, K = 10, 200
B = int(os.environ.get("AB_B", 75776))
lib = _lib.load()
p = {k: v.to(DEV) for k, v in seeded_unet(d, [256, 128, 64], 31).items()}
unet = networks.FullyConnectedUNet(d, (256, 128, 64), 1.0).to(DEV); unet.load_state_dict(p)
udesc, keep = networks.unet_desc(unet)
g = torch.Generator(DEV).manual_seed(1)
states = torch.randn(K + 1, B, d, device=DEV, generator=g)
ts = torch.linspace(0, 1, K + 1, device=DEV)
ldt = ((K + 1) * d + 3) // 4 * 4
target = torch.randn(B, ldt, device=DEV, generator=g)
w = torch.exp(0.5 * torch.randn(B, device=DEV, generator=g))
st = _lib.Setting()
eye, kap = torch.eye(d, device=DEV), torch.ones(d, device=DEV)
st.kind, st.d, st.sigma_is_identity, st.lmbd = 2, d, 1, 1.0
st.sigma, st.sigma_inv, st.kappa, st.nu = eye.data_ptr(), eye.data_ptr(), kap.data_ptr(), kap.data_ptr()
ws = torch.zeros(int(lib.socm_loss_workspace_bytes(udesc, B, K)) // 4 + 1024, device=DEV)
out = {}
for name, flag in (("ffma", _lib.LOSS_FORCE_FFMA), ("tc", _lib.LOSS_FORCE_TC)):
    G = torch.zeros(B, ldt, device=DEV)
    grad = torch.zeros(int(lib.socm_unet_param_count(udesc)), device=DEV)
    loss = torch.zeros(1, device=DEV, dtype=torch.float64)
    _lib.check(lib.socm_unet_loss_fwdbwd_f32(st, udesc, None, ts.data_ptr(), states.data_ptr(), target.data_ptr(), ldt,
                                             w.data_ptr(), None, 1.0 / ((K + 1) * B), B, K, G.data_ptr(), grad.data_ptr(),
                                             loss.data_ptr(), ws.data_ptr(), flag, _lib.stream_ptr()))
    torch.cuda.synchronize()
    out[name] = (float(loss), grad.clone(), G.clone())
nout = [256, 128, 64, d, 256, 128, 128, 256, d]
nin = [d + 1, 256, 128, d + 1, 256, 128, 64, 128, 256]
names = ["down_0", "down_1", "down_2", "res_0", "res_1", "res_2", "up_2", "up_1", "up_0"]
print("loss rel", abs(out["tc"][0] - out["ffma"][0]) / abs(out["ffma"][0]),
      " G rel", float((out["tc"][2] - out["ffma"][2]).norm() / out["ffma"][2].norm()))
off = 0
for n, o, i in zip(names, nout, nin):
    for part, cnt in (("w", o * i), ("b", o)):
        a, b = out["tc"][1][off:off + cnt], out["ffma"][1][off:off + cnt]
        print(f"  {n}.{part}: {float((a - b).norm() / b.norm()):.2e}", end="")
        off += cnt
    print()
