"""Phase cycle counters of the fp16 rollout kernel (build with SOCM_NVCC_EXTRA=-DSOCM_H_PROF into a separate .so):
SOCM_B200_LIB=/path/to/prof.so python scripts/h_prof.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import soc_matching_b200 as sb
from soc_matching_b200 import simulate, _lib
B, K, d = int(os.environ.get("B", 75776)), 200, 10
dev = "cuda"
torch.manual_seed(0)
x0, sigma, sde = sb.make_benchmark_sde("double_well", d, device=dev, gamma=6.0, scaling_factor_M=0.1)
ts = torch.linspace(0, 1, K + 1, device=dev)
simulate.ENGINE = "f16"
w = simulate.rollout(sde, x0.repeat(B, 1), ts, 1.0, seed=5)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * 64)()
lib.socm_debug_h_prof.argtypes = [ctypes.c_void_p]
assert lib.socm_debug_h_prof(buf) == 0
tiles = (B // 128 + 295) // 296
steps = K * tiles
E = ["xin", "r1 chunks (pieces)", "wait D1", "r2", "wait D2", "r3", "wait D3", "o2", "noise", "y1 chunks (pieces)", "wait Y0",
     "y0 / exchange write", "e_sync 1", "res_0 + sde + stores", "e_sync 2", "-"]
M = ["loop", "wait XIN", "d0 pieces + down_1", "wait R2H", "down_2 + res_2 b issue", "wait R3", "res_2 a issue", "wait res_2 done",
     "up_2 issue", "wait O2", "up_1 + up_0 (incl. waits)", "-", "-", "-", "-", "-"]
for name, base, lab in (("owner thread 0", 0, E), ("helper thread 128", 16, E), ("M warp", 32, M)):
    print(name, "(cycles per step):")
    tot = 0
    for i in range(16):
        v = buf[base + i] / steps
        tot += v
        if lab[i] != "-":
            print(f"  {lab[i]:32s} {v:9.0f}")
    print(f"  total {tot:.0f}")

print("owner thread 0, inside the piece -> chunk loops (cycles per piece, 16 pieces per step):")
for i, n in enumerate(["wait PC_FULL", "tcgen05.ld + wait", "fence + arrive PC_EMPTY", "bias/relu/split", "wait CH_EMPTY", "st.shared + fence.proxy.async", "arrive CH_FULL"]):
    print(f"  {n:32s} {buf[48 + i] / steps / 16:9.0f}")
