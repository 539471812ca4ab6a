"""Throughput of the B200 path on the shapes of all five BASELINE.json configs (north star: "throughput on synthetic
batches of each setting's shape"): rollout trajectory-steps/s and SOCM iterations/s at the reference's batch size and
at a large batch, next to the CPU oracle (the reference algorithm) on the same shape at the reference batch size.
Writes one JSON line per case; run on a GPU box:  python scripts/settings_bench.py > gpurun_out/settings_bench.jsonl"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_product_sde, orc, random_setting, seeded_mnet, seeded_unet
import soc_matching_b200 as sb
from soc_matching_b200 import simulate
DEV = "cuda"
CASES = [  # name, kind, d, K, B_ref, hdims_M, algorithms, stopping, warm
    ("C1 OU_quadratic_easy", "ou_quadratic", 20, 50, 128, [128, 128], ["SOCM"], False, False),
    ("C2 OU_linear", "ou_linear", 10, 100, 64, [128, 128], ["SOCM", "SOCM_const_M"], False, False),
    ("C3 OU_quadratic_hard warm start", "ou_quadratic", 20, 150, 64, [128, 128], ["SOCM"], False, True),
    ("C4 molecular_dynamics stopping", "molecular_dynamics", 1, 150, 64, [64, 64], ["SOCM"], True, False),
    ("C5 double_well", "double_well", 10, 200, 128, [128, 128], ["SOCM"], False, False),
]
BIG = int(os.environ.get("BIG", 65536))
CPU = os.environ.get("CPU", "1") == "1"


def ev_time(fn, n):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


for name, kind, d, K, Bref, hm, algos, stopping, warm in CASES:
    st = random_setting(kind, d, seed=3)
    hd = [256, 128, 64]
    unet, mnet = seeded_unet(d, hd, 5), seeded_mnet(d, hm, 6, 0.1, 3 if stopping else 2)
    gam = {"gamma": torch.tensor([2.0 if kind != "double_well" else 6.0]), "gamma2": torch.tensor([1.0]),
           "gamma3": torch.tensor([1.0])}
    wt = None
    if warm:   # an affine warm-start table of the reference's form (models.py:163-199 tabulated), small coefficients
        g = torch.Generator().manual_seed(1)
        wt = orc.WarmStartTable(0.05 * torch.randn(K, d, d, generator=g), 0.05 * torch.randn(K, d, generator=g),
                                0.05 * torch.randn(K + 1, d, d, generator=g), 0.05 * torch.randn(K + 1, d, generator=g))
    x0 = -torch.ones(d) if kind == "molecular_dynamics" else (torch.zeros(d) if kind != "ou_quadratic" else 0.5 * torch.ones(d))
    ts = torch.linspace(0, 1.0, K + 1)
    for B in (Bref, BIG):
        if stopping and B > 65536:
            continue
        sde = make_product_sde(st, unet, mnet, gam, hd, hm, DEV, stopping=stopping, warm=wt)
        xb = x0.to(DEV).repeat(B, 1)
        t_roll = ev_time(lambda: simulate.rollout(sde, xb, ts.to(DEV), st.lmbd), 5 if B == Bref else 3)
        row = {"case": name, "d": d, "K": K, "B": B, "rollout_traj_steps_per_s": B * K / t_roll,
               "rollout_ms": t_roll * 1e3}
        for algo in algos:
            solver = sb.SOC_Solver(sde, x0.to(DEV), None, T=1.0, num_steps=K, lmbd=st.lmbd, d=d, sigma=sde.sigma)

            def it():
                for p in sde.parameters():
                    p.grad = None
                out = solver.loss(B, algorithm=algo, u_warm_start=sde.u_warm_start if warm else None,
                                  use_warm_start=warm, use_stopping_time=stopping)
                out[0].backward()
            t_it = ev_time(it, 5 if B == Bref else 3)
            row[f"{algo}_iters_per_s"] = 1.0 / t_it
            row[f"{algo}_traj_steps_per_s"] = B * K / t_it
        if B == Bref and CPU:   # the reference algorithm (oracle port) on the host cores, same shape
            t0 = time.perf_counter()
            traj = orc.rollout(st, unet, x0.repeat(B, 1), ts, warm=wt)
            t_cpu_roll = time.perf_counter() - t0
            pu = {k: v.clone().requires_grad_(True) for k, v in unet.items()}
            pm = {k: v.clone().requires_grad_(True) for k, v in mnet.items()}
            pg = {k: v.clone().requires_grad_(True) for k, v in gam.items()}
            t0 = time.perf_counter()
            obj, _, _ = orc.socm_loss(st, pu, pm, pg, ts, traj, algorithm=algos[0], warm=wt, use_stopping_time=stopping)
            obj.backward()
            t_cpu_loss = time.perf_counter() - t0
            row["cpu_rollout_traj_steps_per_s"] = B * K / t_cpu_roll
            row[f"cpu_{algos[0]}_iters_per_s"] = 1.0 / (t_cpu_roll + t_cpu_loss)
            row["cpu_threads"] = torch.get_num_threads()
        print(json.dumps(row), flush=True)
