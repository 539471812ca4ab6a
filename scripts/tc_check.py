"""Quick GPU check of the tcgen05 rollout against the FFMA tile kernel (same injected noise)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import make_product_sde, random_setting, seeded_mnet, seeded_unet, rel_l2
from soc_matching_b200 import simulate

DEV = "cuda"
def run(kind, d, K, B, bench=False):
    st = random_setting(kind, d, seed=3)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(st, seeded_unet(d, [256, 128, 64], 5), seeded_mnet(d, [128, 128], 6, 0.1, 3 if kind == "molecular_dynamics" else 2), gam,
                           [256, 128, 64], [128, 128], DEV, stopping=(kind == "molecular_dynamics"))
    ts = torch.linspace(0, 1.0, K + 1, device=DEV)
    x0 = (-torch.ones(B, d, device=DEV)) if kind == "molecular_dynamics" else torch.zeros(B, d, device=DEV)
    noises = torch.randn(K, B, d, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    a = simulate.rollout(sde, x0, ts, 1.0, noises=noises, force_ffma=True)
    torch.cuda.synchronize()
    b = simulate.rollout(sde, x0, ts, 1.0, noises=noises)
    torch.cuda.synchronize()
    print(f"{kind} d={d} K={K} B={B}: states {rel_l2(b.states, a.states):.2e} controls {rel_l2(b.controls, a.controls):.2e} "
          f"lw {rel_l2(b.lw, a.lw):.2e} stop_mismatch {int((a.stop != b.stop).sum())} "
          f"max|ctrl diff| {float((a.controls - b.controls).abs().max()):.2e}", flush=True)
    if bench:
        for name, kw in (("ffma", {"force_ffma": True}), ("tc", {})):
            simulate.rollout(sde, x0, ts, 1.0, seed=1, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                simulate.rollout(sde, x0, ts, 1.0, seed=1, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"   {name}: {ms:.2f} ms  {B * K / ms * 1e3:.3e} traj-steps/s", flush=True)

run("double_well", 10, 3, 128)
run("double_well", 10, 40, 300)
run("ou_quadratic", 20, 25, 70)
run("ou_quadratic", 5, 20, 64)
run("molecular_dynamics", 1, 150, 4096)
run("double_well", 10, 200, 148 * 128 * 4, bench=True)
