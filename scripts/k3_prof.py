"""Per-phase cycle profile of K3a (build with SOCM_NVCC_EXTRA=-DSOCM_TC_PROF)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import seeded_unet
from soc_matching_b200 import _lib, networks
DEV = "cuda"
d, K, B = 10, 200, 148 * 128 * 2 // 201 * 1 + 256
B = 4096
lib = _lib.load()
p = {k: v.to(DEV) for k, v in seeded_unet(d, [256, 128, 64], 31).items()}
unet = networks.FullyConnectedUNet(d, (256, 128, 64), 1.0).to(DEV); unet.load_state_dict(p)
udesc, keep = networks.unet_desc(unet)
g = torch.Generator(DEV).manual_seed(1)
states = torch.randn(K + 1, B, d, device=DEV, generator=g)
ts = torch.linspace(0, 1, K + 1, device=DEV)
ldt = ((K + 1) * d + 3) // 4 * 4
target = torch.randn(B, ldt, device=DEV, generator=g)
w = torch.ones(B, device=DEV)
G = torch.zeros(B, ldt, device=DEV)
grad = torch.zeros(int(lib.socm_unet_param_count(udesc)), device=DEV)
loss = torch.zeros(1, device=DEV, dtype=torch.float64)
ws = torch.zeros(int(lib.socm_loss_workspace_bytes(udesc, B, K)) // 4 + 1024, device=DEV)
st = _lib.Setting()
eye, kap = torch.eye(d, device=DEV), torch.ones(d, device=DEV)
st.kind, st.d, st.sigma_is_identity, st.lmbd = 2, d, 1, 1.0
st.sigma, st.sigma_inv, st.kappa, st.nu = eye.data_ptr(), eye.data_ptr(), kap.data_ptr(), kap.data_ptr()
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.socm_unet_loss_fwdbwd_f32(st, udesc, None, ts.data_ptr(), states.data_ptr(), target.data_ptr(), ldt,
                                             w.data_ptr(), None, 1.0, B, K, G.data_ptr(), grad.data_ptr(), loss.data_ptr(),
                                             ws.data_ptr(), _lib.LOSS_FORCE_TC, _lib.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    n_tiles = (K + 1) * (B // 128)
    print(f"K3 total {e0.elapsed_time(e1):.2f} ms for {n_tiles} tiles -> {e0.elapsed_time(e1) * 1e3 / (n_tiles / 148):.1f} us per tile per SM")
if hasattr(lib, "socm_debug_k3_prof"):
    buf = (ctypes.c_ulonglong * 192)()
    lib.socm_debug_k3_prof(buf)
    tiles = (n_tiles + 147) // 148   # tiles of block 0 in the LAST launch (approximately)
    last = n_tiles % 8192 or 8192
    tiles = (last + 147) // 148
    for who, base in (("owner", 0), ("helper", 64)):
        vals = [buf[base + i] / tiles for i in range(64)]
        print(who, "cycles per tile by phase slot (even = work, odd = wait):")
        print("  " + " ".join(f"{v:.0f}" for v in vals if v > 0))
        print("  total", sum(vals), " work", sum(vals[0::2]), " wait", sum(vals[1::2]))
