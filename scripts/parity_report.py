"""Parity report of the B200 path at the five BASELINE.json shapes (and the 65 536-point reference fixture): prints the
norm-wise relative error of every output / gradient tensor against the CPU oracle, for the default dispatch and for the
forced tcgen05 K3.  Diagnostic companion of tests/test_gpu_baseline_configs.py (which asserts the north-star bounds);
run on a GPU box:  python scripts/parity_report.py [c1 c2 ...] > gpurun_out/parity_report.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import Golden, golden_names, make_product_sde, orc, rel_l2, seeded_mnet, seeded_unet  # noqa: E402
from test_gpu_baseline_configs import HD, _product_run, baseline_config, kink_free_noise  # noqa: E402

DEV = "cuda"


def report(name, kink_free=True):
    st, x0, K, B, hm, gamma, sf_v, algos, stopping, warm = baseline_config(name)
    d = st.d
    unet = seeded_unet(d, HD, 100 + d, sf_v)
    mnet = seeded_mnet(d, hm, 101 + d, 0.1, 3 if stopping else 2)
    gam = {"gamma": torch.tensor([gamma]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    ts = torch.linspace(0, 1.0, K + 1)
    torch.set_num_threads(max(8, torch.get_num_threads()))
    n_redrawn = 0
    if kink_free:
        noises, n_redrawn = kink_free_noise(st, unet, x0, ts, B, warm)
    else:
        import numpy as np
        noises = orc.path_noise(17, np.zeros(B, dtype=np.int64), K, d)
    traj = orc.rollout(st, unet, x0.repeat(B, 1), ts, noises=noises, warm=warm)
    for algo in algos:
        pu = {k: v.clone().requires_grad_(True) for k, v in unet.items()}
        pm = {k: v.clone().requires_grad_(True) for k, v in mnet.items()}
        pg = {k: v.clone().requires_grad_(True) for k, v in gam.items()}
        obj, wm, _ = orc.socm_loss(st, pu, pm, pg, ts, traj, algorithm=algo, warm=warm, use_stopping_time=stopping)
        obj.backward()
        want = {"unet/" + k: v.grad for k, v in pu.items()}
        if algo == "SOCM":
            want.update({"mnet/" + k: v.grad for k, v in pm.items()})
            want["gam/gamma"] = pg["gamma"].grad
        for force_tc in (False, True):
            _, out, grads = _product_run(st, unet, mnet, gam, hm, x0, K, B, noises, algo, stopping, warm, force_tc)
            errs = {k: rel_l2(grads[k].detach().cpu(), w) for k, w in want.items() if grads.get(k) is not None}
            worst = max(errs, key=errs.get)
            print(f"{name} {algo:13s} {'force_tc' if force_tc else 'default ':8s} kink_free={kink_free} "
                  f"(redrawn {n_redrawn}) loss {abs(float(out[0]) - float(obj)) / abs(float(obj)):.1e}  "
                  f"worst grad {errs[worst]:.1e} ({worst})  unet max "
                  f"{max(v for k, v in errs.items() if k.startswith('unet/')):.1e}", flush=True)
            for k, v in errs.items():
                if v > 5e-5:
                    print(f"      {k}: {v:.2e}")


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c2", "c3", "c4", "c5"]
    for n in names:
        for kf in ((False, True) if "--both" in sys.argv else (True,)):
            report(n, kf)
