"""Rollout timing, 3xTF32 engine vs fp16-split engine: python scripts/roll_time.py [B] [K] [d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import soc_matching_b200 as sb
from soc_matching_b200 import simulate
from helpers import rel_l2
B = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
d = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = "cuda"
torch.manual_seed(0)
x0, sigma, sde = sb.make_benchmark_sde("double_well", d, device=dev, gamma=6.0, scaling_factor_M=0.1)
ts = torch.linspace(0, 1, K + 1, device=dev)
xb = x0.repeat(B, 1)
res = {}
for eng in ("tf32", "f16", "ffma"):
    simulate.ENGINE = None if eng == "ffma" else eng
    kw = dict(force_ffma=True) if eng == "ffma" else {}
    if eng == "ffma" and B > 20000:
        continue
    for _ in range(2):
        w = simulate.rollout(sde, xb, ts, 1.0, seed=5, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 5
    for _ in range(n):
        w = simulate.rollout(sde, xb, ts, 1.0, seed=5, **kw)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    res[eng] = (w.states.clone(), w.lw.clone(), w.controls.clone())
    print(f"{eng}: {ms:.3f} ms  {B*K/ms/1e3:.3e} traj-steps/s  finite={bool(torch.isfinite(w.states).all())}", flush=True)
ref = res.get("ffma", res["tf32"])
for eng in res:
    print(eng, "vs", "ffma" if "ffma" in res else "tf32", "states %.2e controls %.2e logw %.2e" % (
        rel_l2(res[eng][0], ref[0]), rel_l2(res[eng][2], ref[2]), rel_l2(res[eng][1], ref[1])))
