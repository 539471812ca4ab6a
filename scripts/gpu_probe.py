"""Quick device-side timing probe (CUDA events) of the individual kernels; scratch tool for
development, prints one line per measurement.  Not the benchmark (see bench.py)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import soc_matching_b200 as sb  # noqa: E402
from soc_matching_b200 import simulate  # noqa: E402


def timed(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    dev = "cuda"
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    torch.manual_seed(0)
    d, K = 10, 200
    x0, sigma, sde = sb.make_benchmark_sde("double_well", d, device=dev, gamma=6.0)
    ts = torch.linspace(0, 1.0, K + 1, device=dev)
    for B in (128, 148 * 64, 1 << 17):
        st0 = x0.repeat(B, 1)
        desc = sb.describe_setting(sde, dev)
        ws = simulate.RolloutWorkspace(desc, sde.nabla_V, B, K, dev, True)
        best, med = timed(lambda: simulate.rollout(sde, st0, ts, 1.0, seed=1, desc=desc, workspace=ws))
        flops = B * K * 338652.0
        print(f"rollout tiled  B={B:7d} K={K}: {best*1e3:9.3f} ms  {B*K/best:.3e} traj-steps/s  {flops/best/1e12:.2f} TFLOP/s")
        if B <= 148 * 64:
            best, med = timed(lambda: simulate.rollout(sde, st0, ts, 1.0, seed=1, desc=desc, workspace=ws,
                                                       force_generic=True), warm=1, reps=2)
            print(f"rollout generic B={B:7d} K={K}: {best*1e3:9.3f} ms  {B*K/best:.3e} traj-steps/s")
    solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sigma)
    for B in (128, 1024, 148 * 64, 1 << 16):
        def it():
            for p in sde.parameters():
                p.grad = None
            out = solver.loss(B, algorithm="SOCM")
            out[0].backward()
        t0 = time.time()
        best, med = timed(it, warm=1, reps=2)
        print(f"SOCM iteration (generic K3) B={B}: {best*1e3:.1f} ms  -> {1/best:.2f} it/s   wall {time.time()-t0:.1f}s")


if __name__ == "__main__":
    main()
