#!/usr/bin/env python
"""Benchmark of the SOC-matching hot path on B200 (contract: see the task's bench.py section).

A "step" is one full SOCM iteration on BASELINE.json's headline configuration
(double_well d=10, num_steps=200, gamma=6, 2^20 synthetic trajectories in total, sharded over the
ranks): Euler-Maruyama rollout -> SOCM target -> importance-weighted loss + backward
(+ one gradient all-reduce when N > 1), through the public API (SOC_Solver.loss / backward).
metric = trajectory-steps/s = B_global * num_steps / t_step.

    python bench.py [--gpus N --steps K --warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

``--impl reference`` times the reference algorithm on the host CPU cores (the oracle port of
oracle/socm_oracle.py -- the reference is pure Python and cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, K_STEPS, GAMMA = 10, 200, 6.0
FLOP_FWD = 338652.0                       # one UNet evaluation at d=10
FLOP_K3_POINT = 338652.0 + 338652.0 + 332800.0   # forward + wgrad + dgrad (no dgrad into [t,x])
# FLOP the kernels actually issue per point (x3 MMAs each): res_1 is folded into up_0, so the
# K = N = 256 GEMM of the reference never runs; small layers are padded to the MMA shapes (K = 16, N = 16 / 32)
EXEC_FWD = 2.0 * (16 * 256 + 256 * 128 + 256 * 16 + 128 * 64 + 64 * 128 + 128 * 128 + 128 * 256 + 256 * 16)
EXEC_DGRAD = 2.0 * (16 * 256 + 256 * 128 + 128 * 128 + 128 * 64 + 64 * 128 + 128 * 256 + 16 * 256)
EXEC_WGRAD = 2.0 * 128 * 1184     # 1184 accumulator columns of 128 rows (csrc/wgrad_tc.cu pass table)
TF32_MEASURED = 1034.0                    # TFLOP/s, all 148 SMs issuing 128x256x8 kind::tf32 MMAs (profiles/r1_umma_probe.log)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region.  The sampler is started
    before the warm-up (nvidia-smi needs a few hundred ms for its first line) and ``stop(t0, t1)`` keeps the samples
    whose timestamp lies inside the timed region; if that region was too short to contain one, the samples taken
    under the warm-up load are used instead and the fact is noted."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/socm_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    @staticmethod
    def _epoch(stamp: str):
        import datetime
        try:
            return datetime.datetime.strptime(stamp.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                rows.append((self._epoch(parts[0]), float(parts[1]), float(parts[2]),
                             [n for n, v in zip(names, parts[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if t0 is not None and r[0] is not None and t0 <= r[0] <= t1]
        used = inside or rows
        out = {"sm_mhz": statistics.median([r[1] for r in used]) if used else None,
               "sm_max_mhz": max(r[2] for r in used) if used else None, "samples": len(used),
               "reasons": sorted({n for r in used for n in r[3]})}
        if used and not inside:
            out["note"] = "timed region shorter than the sampling period: samples of the warm-up load"
        return out


def build_problem(dev, batch_local):
    import soc_matching_b200 as sb
    torch.manual_seed(0)                                             # main.py:71
    x0, sigma, sde = sb.make_benchmark_sde("double_well", D, device=dev, gamma=GAMMA, scaling_factor_M=0.1)
    solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=K_STEPS, lmbd=1.0, d=D, sigma=sigma)
    return sb, sde, solver


def zero_grads(sde):
    for p in sde.parameters():
        p.grad = None


def gpu_arm(args):
    from soc_matching_b200 import dist as sdist
    rank, world, local = sdist.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch
    lo, hi = sdist.shard_bounds(B, rank, world)
    sb, sde, solver = build_problem(dev, hi - lo)
    peaks, peak_src = load_peaks()

    def step():
        zero_grads(sde)
        return sdist.sharded_loss_backward(solver, B, "SOCM")

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    # ---- device-timed region: exactly K steps, inputs resident in HBM
    solver.kernel_events = {}
    barrier()
    wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop(wall0, time.time())
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in solver.kernel_events.items()}
    kernel_n = {k: len(v) for k, v in solver.kernel_events.items()}
    solver.kernel_events = None
    launches = solver.launch_count * args.steps
    if world > 1:
        tt = torch.tensor([t_dev], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t_dev = float(tt)
    ms_per_step = t_dev / args.steps * 1e3
    value = B * K_STEPS / (t_dev / args.steps)

    # ---- end-to-end: same call, inputs from pinned host memory every step, result read back
    x0_host = solver.x0.detach().cpu().pin_memory()
    ts_host = solver.ts.detach().cpu().pin_memory()
    res_host = torch.empty(3, dtype=torch.float32).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        solver.x0 = x0_host.to(dev, non_blocking=True)
        solver.ts = ts_host.to(dev, non_blocking=True)
        val, mw, sw = step()
        res_host.copy_(torch.stack([val.float(), mw, sw]), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(res_host[0])

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        loss_val = e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t_e2e = float(tt)
    e2e_value = B * K_STEPS / (t_e2e / e2e_steps)

    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    # ---- roofline of the dominant kernels (K3 = loss_tc_kernel + wgrad_tc_kernel: UNet forward + loss +
    # dgrad + wgrad at every trajectory point).  achieved = algorithmic FLOP of one K3 call / its
    # CUDA-event time; peak = measured dense bf16 GEMM (sustained, the kernel runs inside a long step).
    # The kernels issue kind::tf32 MMAs three times per product (3xTF32, fp32-class accuracy), so the
    # ceiling of this arithmetic is peak / 2 (tf32 rate) / 3 = peak / 6; both fractions are reported.
    chunk = min(hi - lo, solver.chunk_paths or (1 << 16))
    k3_ms = kernel_ms.get("loss_fwdbwd", float("nan"))
    k3_flop = FLOP_K3_POINT * (K_STEPS + 1) * chunk
    achieved = k3_flop / (k3_ms * 1e-3) / 1e12
    tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1f_k3_traffic.json")
    if os.path.exists(tpath):   # dram bytes of K3a + K3b from the committed ncu capture, per trajectory point
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_per_point"] * (K_STEPS + 1) * chunk
    k1_ms = kernel_ms.get("rollout", float("nan"))
    k1_tflops = FLOP_FWD * K_STEPS * chunk / (k1_ms * 1e-3) / 1e12
    roofline = {
        "kernel": "K3: loss_tc_kernel + wgrad_tc_kernel (tcgen05, 3xTF32)", "bound": "tensor",
        "achieved": round(achieved, 2), "peak": tensor_peak, "unit": "TFLOP/s", "frac": round(achieved / tensor_peak, 4),
        "traffic": traffic, "peak_source": f"{peak_src} bf16_tflops_sustained",
        "executed_tflops": round(achieved * (EXEC_FWD + EXEC_DGRAD + EXEC_WGRAD) / FLOP_K3_POINT, 2),
        "frac_of_3xtf32_ceiling": round(achieved * (EXEC_FWD + EXEC_DGRAD + EXEC_WGRAD) / FLOP_K3_POINT
                                        / (TF32_MEASURED / 3.0), 4),
        "note": "achieved = ALGORITHMIC fp32 FLOP of the reference's network (SURVEY.md 8d) / time; the kernels issue "
                "fewer (executed_tflops): res_1 is folded into up_0 algebraically (DESIGN.md 3.1).  Every product is "
                "3 kind::tf32 MMAs, so the ceiling of the executed arithmetic is the measured dense tf32 rate "
                f"({TF32_MEASURED:.0f} TFLOP/s, scripts/umma_probe.cu, profiles/r1_umma_probe.log) / 3 "
                "(frac_of_3xtf32_ceiling uses the executed FLOP); DRAM traffic is dominated by the wgrad operand "
                "scratch (DESIGN.md 3.4)",
        "algorithmic_flop_per_launch": k3_flop, "avg_launch_ms": round(k3_ms, 3),
        "kernel_share_of_step": round(k3_ms * kernel_n.get("loss_fwdbwd", 0) / args.steps / ms_per_step, 3),
        "rollout": {"kernel": "rollout_tc_kernel (tcgen05, 3xTF32)", "achieved_tflops": round(k1_tflops, 2),
                    "frac": round(k1_tflops / tensor_peak, 4),
                    "executed_tflops": round(k1_tflops * EXEC_FWD / FLOP_FWD, 2),
                    "frac_of_3xtf32_ceiling": round(k1_tflops * EXEC_FWD / FLOP_FWD / (TF32_MEASURED / 3.0), 4),
                    "hbm_gbs": round(chunk * K_STEPS * 128 / (k1_ms * 1e-3) / 1e9, 1),
                    "hbm_frac": round(chunk * K_STEPS * 128 / (k1_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]), 4)},
    }
    cpu = cpu_baseline(sample_batch=128, reps=1) if world == 1 and not args.no_cpu else None
    line = {
        "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "double_well d=10 num_steps=200 gamma=6 SOCM (rollout + target + loss + backward)",
                   "global_batch": B, "paths_per_gpu": hi - lo, "chunk_paths": chunk, "hdims": [256, 128, 64],
                   "hdims_M": [128, 128], "noise": "in-kernel Philox4x32-10", "arithmetic": "fp32 (UNet GEMMs as 3xTF32 on tcgen05)",
                   "l2": "working set per step (GBs of trajectories) is far larger than the 126 MB L2"},
        "socm_iters_per_s": 1e3 / ms_per_step,
        "kernel_ms_avg_per_launch": {k: round(v, 3) for k, v in kernel_ms.items()},
        "kernel_launches_per_step": {k: v // args.steps for k, v in kernel_n.items()},
        "rollout_traj_steps_per_s": chunk * K_STEPS / (kernel_ms["rollout"] * 1e-3) if "rollout" in kernel_ms else None,
        "e2e": {"value": e2e_value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": (D + K_STEPS + 1) * 4,
                "d2h_bytes_per_step": 12, "steps": e2e_steps, "loss": loss_val},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def cpu_baseline(sample_batch=128, reps=1):
    """The reference algorithm (oracle port: Python loop rollout, jacrev, 5-D einsums, autograd
    backward) on the host cores, on a bounded sample of the same workload."""
    from oracle import socm_oracle as orc
    torch.manual_seed(0)
    kappa, nu = torch.ones(D), torch.ones(D)
    kappa[:3], nu[:3] = 5, 3
    st = orc.Setting("double_well", D, torch.eye(D), 1.0, kappa=kappa, nu=nu)
    import soc_matching_b200.networks as nets  # parameter containers only (CPU tensors)
    unet_m = nets.FullyConnectedUNet(D, (256, 128, 64), 1.0)
    mnet_m = nets.SigmoidMLP(D, (128, 128), torch.nn.Parameter(torch.tensor([GAMMA])), 0.1)
    unet = {k: v for k, v in unet_m.named_parameters()}
    mnet = {k: v for k, v in mnet_m.named_parameters() if k.startswith("sigmoid_layers")}
    gam = {"gamma": mnet_m.gamma}
    ts = torch.linspace(0, 1.0, K_STEPS + 1)
    x0 = torch.zeros(sample_batch, D)
    times = []
    for _ in range(reps):
        for p in list(unet.values()) + list(mnet.values()) + [gam["gamma"]]:
            p.grad = None
        t0 = time.perf_counter()
        traj = orc.rollout(st, unet, x0, ts)
        obj, _, _ = orc.socm_loss(st, unet, mnet, gam, ts, traj, algorithm="SOCM")
        obj.backward()
        times.append(time.perf_counter() - t0)
    t = min(times)
    return {"value": sample_batch * K_STEPS / t, "unit": "trajectory-steps/s", "cores": torch.get_num_threads(),
            "kind": "port", "seconds_per_iteration": t,
            "sample": f"one SOCM iteration (rollout + loss + backward) at B={sample_batch} of the same "
                      f"double_well d=10 K=200 workload, torch {torch.__version__} CPU fp32"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for multi-process launches: give the CPU reference the host cores it can use
    # (the same count torch picks by default in a single-process run on this box: the affinity mask, at most 16)
    try:
        torch.set_num_threads(max(1, min(16, len(os.sched_getaffinity(0)))))
    except (AttributeError, OSError):
        pass
    B = 128
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_baseline(B, 1)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(B, 1)
    t = (time.perf_counter() - t0) / args.steps
    value = B * K_STEPS / t
    last.update(value=value, seconds_per_iteration=t)
    print(json.dumps({
        "impl": "reference", "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "double_well d=10 num_steps=200 gamma=6 SOCM (rollout + target + loss + backward)",
                   "sample_batch": B},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1 << 20, help="global number of trajectories")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
