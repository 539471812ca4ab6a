"""Where does a small-batch SOCM iteration spend its time?  C5 at the reference's batch size (double_well d=10, K=200,
B=128): wall clock vs device time of the rollout alone, loss(), loss()+backward(), and the per-kernel CUDA-event times
the solver records.  python scripts/small_batch_prof.py [B]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import soc_matching_b200 as sb
from soc_matching_b200 import simulate

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = "cuda"
torch.manual_seed(0)
x0, sigma, sde = sb.make_benchmark_sde("double_well", 10, device=dev, gamma=6.0, scaling_factor_M=0.1)
solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=200, lmbd=1.0, d=10, sigma=sigma)
xb = x0.repeat(B, 1)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, a.elapsed_time(b) / n


def it(backward=True):
    for p in sde.parameters():
        p.grad = None
    out = solver.loss(B, algorithm="SOCM")
    if backward:
        out[0].backward()


print(f"B={B}: rollout          wall %.3f ms  device %.3f ms" % timeit(lambda: simulate.rollout(sde, xb, solver.ts, 1.0)))
print(f"B={B}: loss             wall %.3f ms  device %.3f ms" % timeit(lambda: it(False)))
print(f"B={B}: loss + backward  wall %.3f ms  device %.3f ms" % timeit(lambda: it(True)))
solver.kernel_events = {}
it(True)
torch.cuda.synchronize()
for k, v in solver.kernel_events.items():
    print(f"   {k:12s} {sum(a.elapsed_time(b) for a, b in v):.3f} ms ({len(v)} calls)")
solver.kernel_events = None
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    it(True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
