#!/usr/bin/env python
"""Benchmark of the SOC-matching hot path on B200 (contract: see the task's bench.py section).

A "step" is one full SOCM iteration of a BASELINE.json configuration through the public API (SOC_Solver.loss /
backward): Euler-Maruyama rollout -> SOCM target -> importance-weighted loss + backward (+ one gradient all-reduce when
N > 1).  The default (headline) workload is config 5: double_well d=10, num_steps=200, gamma=6, 2^20 synthetic
trajectories in total, sharded over the ranks; ``--config c1..c4`` runs the other four configurations at their own
shapes (settings.py:215-289) with a synthetic batch (default 65 536, ``--batch`` to change).
metric = trajectory-steps/s = B_global * num_steps / t_step.

    python bench.py [--gpus N --steps K --warmup W] [--config c1..c5] [--batch B] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

``--impl reference`` times the reference algorithm on the host CPU cores (the oracle port of
oracle/socm_oracle.py -- the reference is pure Python, its setup.py installs no code (py_modules names a module that
does not exist), and /root/reference cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TF32_MEASURED = 1034.0                    # TFLOP/s, all 148 SMs issuing 128x256x8 kind::tf32 MMAs (profiles/r1_umma_probe.log)

# BASELINE.json configs: setting, d, K, README batch, hdims_M, gamma, sf_nabla_V, stopping, warm start
CONFIGS = {
    "c1": dict(setting="OU_quadratic_easy", d=20, K=50, ref_batch=128, hdims_M=(128, 128), gamma=2.0, sf_v=1.0),
    "c2": dict(setting="OU_linear", d=10, K=100, ref_batch=64, hdims_M=(128, 128), gamma=2.0, sf_v=1.0),
    "c3": dict(setting="OU_quadratic_hard", d=20, K=150, ref_batch=64, hdims_M=(128, 128), gamma=2.0, sf_v=0.1, warm=True),
    "c4": dict(setting="molecular_dynamics", d=1, K=150, ref_batch=64, hdims_M=(64, 64), gamma=2.0, sf_v=1.0,
               stopping=True),
    "c5": dict(setting="double_well", d=10, K=200, ref_batch=128, hdims_M=(128, 128), gamma=6.0, sf_v=1.0),
}


def flop_model(d: int):
    """Algorithmic FLOP of the reference's network per trajectory point and the FLOP the
    tcgen05 kernels issue (x3 MMAs each): res_1 is folded into up_0, small layers are padded to the
    MMA shapes."""
    fwd = 2.0 * ((d + 1) * 256 + 256 * 128 + 128 * 64 + 64 * 128 + 128 * 128 + 128 * 256 + 256 * 256 + 256 * d + (d + 1) * d)
    dgrad = fwd - 2.0 * ((d + 1) * 256 + (d + 1) * d)          # no gradient into [t, x]
    kin = ((d + 1 + 7) // 8) * 8
    ny = 16 if kin <= 16 else 32
    ex_fwd = 2.0 * (kin * 256 + 256 * 128 + 256 * ny + 128 * 64 + 64 * 128 + 128 * 128 + 128 * 256 + 256 * ny)
    ex_dgrad = 2.0 * (kin * 256 + 256 * 128 + 128 * 128 + 128 * 64 + 64 * 128 + 128 * 256 + kin * 256)
    ex_wgrad = 2.0 * 128 * 1184          # 1184 accumulator columns of 128 rows (csrc/wgrad_tc.cu pass table)
    return dict(fwd=fwd, k3=fwd + fwd + dgrad, ex_fwd=ex_fwd, ex_k3=ex_fwd + ex_dgrad + ex_wgrad)


def k2_flop(d: int, K: int, n_paths: int) -> float:
    """Triangular (s >= t) contraction, forward; the backward is the same again."""
    return 2.0 * d * d * n_paths * K * (K + 1) + 2.0 * d * d * n_paths * (K + 1)


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


def build_problem(dev, cfg):
    """The configuration's problem exactly as settings.define_variables builds it under torch.manual_seed(0)
    (main.py:71); config 3's warm start is an affine table of the form models.py:163-199 tabulates to (the spline fit
    itself is out of scope; the cost per call does not depend on its values)."""
    import soc_matching_b200 as sb
    torch.manual_seed(0)
    d, K = cfg["d"], cfg["K"]
    x0, sigma, sde = sb.make_benchmark_sde(cfg["setting"], d, device=dev, gamma=cfg["gamma"], scaling_factor_M=0.1,
                                           hdims_M=cfg["hdims_M"], scaling_factor_nabla_V=cfg["sf_v"],
                                           use_stopping_time=bool(cfg.get("stopping")))
    if cfg.get("warm"):
        g = torch.Generator().manual_seed(1)
        tt = torch.linspace(0, 1, K + 1)
        eye = torch.eye(d)
        A_l = -(0.5 + tt).reshape(-1, 1, 1) * eye + 0.05 * torch.randn(K + 1, d, d, generator=g)
        c_l = 0.3 * torch.sin(3.0 * tt).reshape(-1, 1) * torch.randn(1, d, generator=g)
        sde.u_warm_start = sb.WarmStartTable(A_l[:-1].clone(), c_l[:-1].clone(), A_l, c_l).to(dev)
        sde.use_warm_start = True
    solver = sb.SOC_Solver(sde, x0, None, T=1.0, num_steps=K, lmbd=1.0, d=d, sigma=sigma)
    return sb, sde, solver


def loss_kwargs(cfg, sde):
    kw = {}
    if cfg.get("warm"):
        kw.update(u_warm_start=sde.u_warm_start, use_warm_start=True)
    if cfg.get("stopping"):
        kw.update(use_stopping_time=True)
    return kw


def zero_grads(solver):
    for p in solver.parameters():
        p.grad = None


def kernel_source_hash():
    """sha256 over the CODE of the K3 kernel sources (// comments and blank lines stripped): a committed ncu traffic figure
    is only reported while it still describes the kernels that are being timed."""
    h = hashlib.sha256()
    for f in ("loss_h.cu", "wgrad_h.cu", "unet_h.cuh", "loss_tc.cuh", "wgrad_tables.cuh", "umma.cuh"):
        with open(os.path.join(ROOT, "soc_matching_b200", "csrc", f), "r") as fh:
            for line in fh:
                code = line.split("//", 1)[0].strip()
                if code:
                    h.update(code.encode() + b"\n")
    return h.hexdigest()[:16]


def multi_rank_check(sdist, cfg, dev, world, rank):
    """N ranks == 1 rank: one small sharded iteration on a fixed Philox key against the same batch on rank 0 alone
    (path-indexed counters make the noise independent of the sharding).  Returns the worst relative gradient error."""
    import soc_matching_b200 as sb
    from soc_matching_b200 import simulate
    Bc = 4096
    _, sde, solver = build_problem(dev, cfg)
    kw = loss_kwargs(cfg, sde)
    simulate._SEED_COUNTER[0] = 10_000
    zero_grads(solver)
    val, _, _ = sdist.sharded_loss_backward(solver, Bc, "SOCM", **kw)
    sharded = {n: p.grad.detach().clone() for n, p in solver.named_parameters() if p.grad is not None}
    out = None
    if rank == 0:
        simulate._SEED_COUNTER[0] = 10_000
        zero_grads(solver)
        solver.path_offset = 0
        res = solver.loss(Bc, algorithm="SOCM", **kw)
        res[0].backward()
        worst = 0.0
        for n, g in sharded.items():
            ref = dict(solver.named_parameters())[n].grad
            den = float(torch.linalg.norm(ref.double()))
            if den > 0:
                worst = max(worst, float(torch.linalg.norm((g - ref).double())) / den)
        out = {"batch": Bc, "loss_rel_err": abs(float(val) - float(res[0])) / abs(float(res[0])),
               "max_grad_rel_err": worst}
    torch.distributed.barrier()
    return out


def gpu_arm(args):
    from soc_matching_b200 import dist as sdist
    cfg = CONFIGS[args.config]
    d, K = cfg["d"], cfg["K"]
    rank, world, local = sdist.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch if args.batch else ((1 << 20) if args.config == "c5" else (1 << 16))
    lo, hi = sdist.shard_bounds(B, rank, world)
    sb, sde, solver = build_problem(dev, cfg)
    kw = loss_kwargs(cfg, sde)
    peaks, peak_src = load_peaks()
    fm = flop_model(d)

    def step():
        zero_grads(solver)
        return sdist.sharded_loss_backward(solver, B, "SOCM", **kw)

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
    # summed over every launch of the timed region (the last chunk of a shard is shorter than the others: an average
    # launch time divided into a full chunk's FLOP would overstate the rate)
    kernel_ms_sum = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in solver.kernel_events.items()}
    kernel_n = {k: len(v) for k, v in solver.kernel_events.items()}
    solver.kernel_events = None
    launches = solver.launch_count * args.steps
    if world > 1:
        tt = torch.tensor([t_dev], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t_dev = float(tt)
    ms_per_step = t_dev / args.steps * 1e3
    value = B * K / (t_dev / args.steps)

    # ---- end-to-end: same call, inputs from pinned host memory every step, result read back; all --steps
    x0_host = solver.x0.detach().cpu().pin_memory()
    ts_host = solver.ts.detach().cpu().pin_memory()
    res_host = torch.empty(3, dtype=torch.float32).pin_memory()

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
    for _ in range(args.steps):
        loss_val = e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t_e2e = float(tt)
    e2e_value = B * K / (t_e2e / args.steps)

    check = multi_rank_check(sdist, cfg, dev, world, rank) if world > 1 else None
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    # ---- rooflines (rank 0's shard).  achieved = ALGORITHMIC fp32 FLOP of all launches of the timed region / the sum
    # of their CUDA-event times; peak = measured dense bf16 GEMM (sustained: the kernels run inside a long step).
    # Every product of the UNet kernels is three tensor-core MMAs (fp32-class accuracy): fp16 hi/lo splits on kind::f16
    # (the engine this workload runs, csrc/unet_h.cuh, target_h.cu) or 3xTF32 (csrc/unet_tc.cuh), so the ceiling of the
    # executed arithmetic is the measured dense rate of that kind / 3; both fractions are reported.
    n_local = hi - lo
    points = (K + 1) * n_local * args.steps
    tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))

    def rate(flop, name):
        ms = kernel_ms_sum.get(name)
        return flop / (ms * 1e-3) / 1e12 if ms else None

    k3 = rate(fm["k3"] * points, "loss_fwdbwd")
    k1 = rate(fm["fwd"] * K * n_local * args.steps, "rollout")
    k2f = rate(k2_flop(d, K, n_local) * args.steps, "target")
    k2b = rate(k2_flop(d, K, n_local) * args.steps, "target_bwd")
    traffic, traffic_src = None, "no ncu capture of the current K3 sources committed"
    tpath = os.path.join(ROOT, "profiles", "r2_k3_traffic.json")
    if os.path.exists(tpath):   # dram bytes of K3a + K3b per trajectory point from one `ncu --set full` capture
        tj = json.load(open(tpath))
        if tj.get("kernel_source_hash") == kernel_source_hash():
            traffic = tj["dram_bytes_per_point"] * (K + 1) * min(n_local, solver.chunk_paths or n_local)
            traffic_src = "profiles/r2_k3_traffic.json (ncu --set full, same kernel sources)"
        else:
            traffic_src = "profiles/r2_k3_traffic.json is stale (kernel sources changed since the capture)"
    chunk = min(n_local, solver.chunk_paths or n_local)

    def frac(x, scale=1.0):
        return None if x is None else round(x * scale / tensor_peak, 4)

    # engine of the UNet kernels, as csrc/loss.cu:95 and csrc/rollout.cu:277 choose it (fp16 split wherever the default
    # architecture with d <= 15 runs on the tensor cores, unless SOCM_F16=0 / simulate.ENGINE say otherwise)
    f16 = d <= 15 and os.environ.get("SOCM_F16") != "0" and sb.simulate.ENGINE != "tf32"
    unet_ceiling = tensor_peak / 3.0 if f16 else TF32_MEASURED / 3.0
    eng = "3 x kind::f16 on fp16 hi/lo splits" if f16 else "3xTF32"
    sfx = "h" if f16 else "tc"
    # the target GEMMs run the fp16-split engine for every d (csrc/target_h.cu, target_bwd_h.cu) unless 3xTF32 is asked for
    f16_k2 = os.environ.get("SOCM_F16") != "0" and sb.simulate.ENGINE != "tf32"
    k2_ceiling = tensor_peak / 3.0 if f16_k2 else TF32_MEASURED / 3.0
    eng2 = "3 x kind::f16 on fp16 hi/lo splits" if f16_k2 else "3xTF32"
    sfx2 = "h" if f16_k2 else "tc"

    def ceil_frac(x, ex_ratio, ceiling=None):
        return None if x is None else round(x * ex_ratio / (ceiling or unet_ceiling), 4)

    rnd = lambda x: None if x is None else round(x, 2)  # noqa: E731
    roofline = {
        "kernel": f"K3: loss_{sfx}_kernel + wgrad_{sfx}_kernel (tcgen05, {eng})", "bound": "tensor",
        "achieved": rnd(k3), "peak": tensor_peak, "unit": "TFLOP/s", "frac": frac(k3),
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": f"{peak_src} bf16_tflops_sustained",
        "executed_tflops": rnd(k3 * fm["ex_k3"] / fm["k3"]) if k3 else None,
        "frac_of_3mma_ceiling": ceil_frac(k3, fm["ex_k3"] / fm["k3"]),
        "ceiling_3mma_tflops": round(unet_ceiling, 1),
        "note": "achieved = ALGORITHMIC fp32 FLOP of the reference's network (SURVEY.md 8d) summed over all launches of "
                "the timed region / summed CUDA-event time; the kernels issue fewer (executed_tflops): res_1 is folded "
                "into up_0 algebraically (DESIGN.md 3.1).  Every product is 3 MMAs, so the ceiling of the executed "
                "arithmetic is a third of the dense rate of the MMA kind: " +
                ("the measured bf16/fp16 peak above (kind::f16 issues at the bf16 rate, profiles/r2_f16_probe.log)" if f16
                 else f"the measured dense tf32 rate ({TF32_MEASURED:.0f} TFLOP/s, profiles/r1_umma_probe.log)") +
                ".  K3a + K3b together are bounded by the HBM traffic of the activation scratch (traffic, DESIGN.md 3.4)",
        "algorithmic_flop_per_launch": fm["k3"] * (K + 1) * chunk,
        "launch_ms_sum": round(kernel_ms_sum.get("loss_fwdbwd", float("nan")), 3),
        "kernel_share_of_step": round(kernel_ms_sum.get("loss_fwdbwd", 0.0) / args.steps / ms_per_step, 3),
        "rollout": {"kernel": f"rollout_{sfx}_kernel (tcgen05, {eng})", "achieved_tflops": rnd(k1), "frac": frac(k1),
                    "frac_of_3mma_ceiling": ceil_frac(k1, fm["ex_fwd"] / fm["fwd"]),
                    "hbm_gbs": rnd(n_local * K * args.steps * (12 * d + 8) / (kernel_ms_sum["rollout"] * 1e-3) / 1e9),
                    "hbm_frac": round(n_local * K * args.steps * (12 * d + 8) / (kernel_ms_sum["rollout"] * 1e-3) / 1e9
                                      / float(peaks["hbm_gbs"]), 4),
                    "share_of_step": round(kernel_ms_sum["rollout"] / args.steps / ms_per_step, 3)},
        "target": {"kernel": (f"target_{sfx2}_kernel / target_bwd_{sfx2}_kernel (tcgen05, {eng2}; grouped SIMT with stopping times)"),
                   "fwd_tflops": rnd(k2f), "fwd_frac": frac(k2f), "fwd_frac_of_3mma_ceiling": ceil_frac(k2f, 1.0, k2_ceiling),
                   "bwd_tflops": rnd(k2b), "bwd_frac": frac(k2b), "bwd_frac_of_3mma_ceiling": ceil_frac(k2b, 1.0, k2_ceiling),
                   "share_of_step": round((kernel_ms_sum.get("target", 0.0) + kernel_ms_sum.get("target_bwd", 0.0))
                                          / args.steps / ms_per_step, 3)},
    }
    cpu = cpu_baseline(args.config, reps=1, warm=True) if world == 1 and not args.no_cpu else None
    line = {
        "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg) + " SOCM (rollout + target + loss + backward)", "name": args.config,
                   "global_batch": B, "paths_per_gpu": n_local, "chunk_paths": chunk, "hdims": [256, 128, 64],
                   "hdims_M": list(cfg["hdims_M"]), "noise": "in-kernel Philox4x32-10",
                   "arithmetic": f"fp32 (UNet GEMMs on tcgen05 as {eng}, fp32 accumulation)",
                   "l2": "working set per step (GBs of trajectories) is far larger than the 126 MB L2"},
        "socm_iters_per_s": 1e3 / ms_per_step,
        "kernel_ms_per_step": {k: round(v / args.steps, 3) for k, v in kernel_ms_sum.items()},
        "kernel_launches_per_step": {k: v // args.steps for k, v in kernel_n.items()},
        "rollout_traj_steps_per_s": n_local * K * args.steps / (kernel_ms_sum["rollout"] * 1e-3),
        "e2e": {"value": e2e_value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": (d + K + 1) * 4,
                "d2h_bytes_per_step": 12, "steps": args.steps, "loss": loss_val},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    if check is not None:
        line["multi_rank_check"] = check
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def workload_name(cfg):
    extra = " warm start" if cfg.get("warm") else (" stopping times" if cfg.get("stopping") else "")
    return f"{cfg['setting']} d={cfg['d']} num_steps={cfg['K']} gamma={cfg['gamma']:g}{extra}"


def cpu_problem(cfg):
    """The same configuration for the CPU oracle (plain tensors)."""
    from oracle import socm_oracle as orc
    import soc_matching_b200.networks as nets  # parameter containers only (CPU tensors)
    torch.manual_seed(0)
    d, K = cfg["d"], cfg["K"]
    eye = torch.eye(d)
    name = cfg["setting"]
    if name.startswith("OU_quadratic"):
        a, p, q = (1.0, 1.0, 0.5) if name.endswith("hard") else (0.2, 0.2, 0.1)
        x0 = 0.5 * torch.randn(d)
        st = orc.Setting("ou_quadratic", d, eye.clone(), 1.0, A=a * eye, P=p * eye, Q=q * eye)
    elif name == "OU_linear":
        x0 = torch.zeros(d)
        xi = 0.1 * torch.randn(d, d)
        st = orc.Setting("ou_linear", d, eye + xi, 1.0, A=-eye + xi, omega=torch.ones(d))
    elif name == "double_well":
        x0 = torch.zeros(d)
        kappa, nu = torch.ones(d), torch.ones(d)
        kappa[:3], nu[:3] = 5, 3
        st = orc.Setting("double_well", d, eye.clone(), 1.0, kappa=kappa, nu=nu)
    else:
        x0 = -torch.ones(d)
        st = orc.Setting("molecular_dynamics", d, eye.clone(), 1.0, kappa=torch.ones(d))
    stopping = bool(cfg.get("stopping"))
    unet_m = nets.FullyConnectedUNet(d, (256, 128, 64), cfg["sf_v"])
    gam = {"gamma": torch.nn.Parameter(torch.tensor([cfg["gamma"]])), "gamma2": torch.nn.Parameter(torch.tensor([1.0])),
           "gamma3": torch.nn.Parameter(torch.tensor([1.0]))}
    if stopping:
        mnet_m = nets.TwoBoundarySigmoidMLP(d, cfg["hdims_M"], gam["gamma"], gam["gamma2"], gam["gamma3"], 0.1)
    else:
        mnet_m = nets.SigmoidMLP(d, cfg["hdims_M"], gam["gamma"], 0.1)
    unet = {k: v for k, v in unet_m.named_parameters()}
    mnet = {k: v for k, v in mnet_m.named_parameters() if k.startswith("sigmoid_layers")}
    warm = None
    if cfg.get("warm"):
        g = torch.Generator().manual_seed(1)
        tt = torch.linspace(0, 1, K + 1)
        A_l = -(0.5 + tt).reshape(-1, 1, 1) * eye + 0.05 * torch.randn(K + 1, d, d, generator=g)
        c_l = 0.3 * torch.sin(3.0 * tt).reshape(-1, 1) * torch.randn(1, d, generator=g)
        warm = orc.WarmStartTable(A_l[:-1].clone(), c_l[:-1].clone(), A_l, c_l)
    return orc, st, x0, unet, mnet, gam, warm, stopping


def cpu_baseline(config="c5", reps=1, warm=True, sample_batch=None):
    """The reference algorithm (oracle port: Python-loop rollout, jacrev, 5-D einsums, autograd backward) on the host
    cores, on a bounded sample of the same workload: one SOCM iteration at the README's batch size of the
    configuration.  ``warm``: one untimed call first (thread pools, allocator)."""
    cfg = CONFIGS[config]
    K = cfg["K"]
    B = sample_batch or cfg["ref_batch"]
    orc, st, x0, unet, mnet, gam, warm_tab, stopping = cpu_problem(cfg)
    ts = torch.linspace(0, 1.0, K + 1)
    xb = x0.repeat(B, 1)
    times = []
    for r in range(reps + (1 if warm else 0)):
        for p in list(unet.values()) + list(mnet.values()) + list(gam.values()):
            p.grad = None
        t0 = time.perf_counter()
        traj = orc.rollout(st, unet, xb, ts, warm=warm_tab)
        obj, _, _ = orc.socm_loss(st, unet, mnet, gam, ts, traj, algorithm="SOCM", warm=warm_tab,
                                  use_stopping_time=stopping)
        obj.backward()
        if r > 0 or not warm:
            times.append(time.perf_counter() - t0)
    t = min(times)
    return {"value": B * K / t, "unit": "trajectory-steps/s", "cores": torch.get_num_threads(),
            "kind": "port", "seconds_per_iteration": t,
            "sample": f"one SOCM iteration (rollout + loss + backward) at B={B} (the README's batch size) of the same "
                      f"{workload_name(cfg)} workload, after one warm-up call, torch {torch.__version__} CPU fp32"}


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
    cfg = CONFIGS[args.config]
    B, K = cfg["ref_batch"], cfg["K"]
    cpu_baseline(args.config, 1, warm=False)          # warm-up
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(args.config, 1, warm=False)
    t = (time.perf_counter() - t0) / args.steps
    value = B * K / t
    last.update(value=value, seconds_per_iteration=t)
    # a second, larger sample so that one ratio is closer to the GPU arm's regime: the rollout alone (the loss at this
    # batch would need a (K+1)^2 B d^2 intermediate of ~1 TB in the reference's formulation)
    orc, st, x0, unet, _, _, warm_tab, _ = cpu_problem(cfg)
    Bl = 16384
    t1 = time.perf_counter()
    orc.rollout(st, unet, x0.repeat(Bl, 1), torch.linspace(0, 1.0, K + 1), warm=warm_tab)
    t_roll = time.perf_counter() - t1
    last["rollout_only_large_batch"] = {"batch": Bl, "seconds": t_roll, "value": Bl * K / t_roll,
                                        "unit": "trajectory-steps/s (rollout only)"}
    print(json.dumps({
        "impl": "reference", "metric": "trajectory-steps/s", "value": value, "unit": "trajectory-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg) + " SOCM (rollout + target + loss + backward)", "name": args.config,
                   "sample_batch": B},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: the headline c5)")
    ap.add_argument("--batch", type=int, default=0, help="global number of trajectories (default 2^20 for c5, 2^16 otherwise)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
