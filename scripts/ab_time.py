"""Kernel timing for A/B runs of development builds (SOCM_B200_LIB=<path> selects the library):
K1 (rollout) and K3 (loss + backward) through the C ABI on the bench workload's shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import seeded_unet
from soc_matching_b200 import _lib, networks, simulate, sde as sde_mod
DEV = "cuda"
d, K = 10, 200
B = int(os.environ.get("AB_B", 8140 * 128 // 201 // 128 * 128 * 2))
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
def k3():
    _lib.check(lib.socm_unet_loss_fwdbwd_f32(st, udesc, None, ts.data_ptr(), states.data_ptr(), target.data_ptr(), ldt,
                                             w.data_ptr(), None, 1.0, B, K, G.data_ptr(), grad.data_ptr(), loss.data_ptr(),
                                             ws.data_ptr(), _lib.LOSS_FORCE_TC, _lib.stream_ptr()))
def timeit(fn, n=4):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t3 = timeit(k3)
n_tiles = (K + 1) * (B // 128)
print(f"{os.environ.get('SOCM_B200_LIB', 'default')}: B={B} K3 {t3:.2f} ms ({n_tiles} tiles, {t3 * 1e3 / (n_tiles / 148):.1f} us/tile/SM)", end="")
if os.environ.get("AB_K2", "1") == "1":
    B2 = 37888
    ldr = ((2 * K + 1) * d + 3) // 4 * 4
    G2 = torch.randn(B2, ldt, device=DEV, generator=g)
    R2 = torch.randn(B2, ldr, device=DEV, generator=g)
    dL = torch.zeros((K + 1) * d, ldr, device=DEV)
    wsb = torch.empty(int(lib.socm_target_gemm_bwd_tc_workspace_bytes(B2, K, d)), device=DEV, dtype=torch.uint8)
    t2 = timeit(lambda: _lib.check(lib.socm_target_gemm_bwd_tc_f32(G2.data_ptr(), R2.data_ptr(), B2, K, d, ldr, ldt,
                                                                    dL.data_ptr(), 1, wsb.data_ptr(), _lib.stream_ptr())), 3)
    print(f"  K2b {t2:.2f} ms for {B2} paths", end="")
if os.environ.get("AB_K2F", "0") == "1":
    B2 = 75776
    ldr = ((2 * K + 1) * d + 3) // 4 * 4
    nrows = (K + 1) * d
    Lm = torch.randn(nrows, ldr, device=DEV, generator=g)
    i_of_row = torch.arange(nrows, device=DEV) // d
    col = torch.arange(ldr, device=DEV)
    Lm[(col[None, :] < 2 * i_of_row[:, None] * d) | (col[None, :] >= (2 * K + 1) * d)] = 0.0
    R2 = torch.randn(B2, ldr, device=DEV, generator=g)
    T2 = torch.empty(B2, ldt, device=DEV)
    wsf = torch.empty(int(lib.socm_target_gemm_tc_workspace_bytes(K, d)), device=DEV, dtype=torch.uint8)
    t2f = timeit(lambda: _lib.check(lib.socm_target_gemm_tc_f32(Lm.data_ptr(), R2.data_ptr(), B2, K, d, ldr, T2.data_ptr(), ldt,
                                                                wsf.data_ptr(), _lib.stream_ptr())), 3)
    print(f"  K2f {t2f:.2f} ms for {B2} paths", end="")
if os.environ.get("AB_K1", "1") == "1":
    from helpers import make_product_sde, random_setting, seeded_mnet
    stg = random_setting("double_well", d, seed=3)
    gam = {"gamma": torch.tensor([6.0]), "gamma2": torch.tensor([1.0]), "gamma3": torch.tensor([1.0])}
    sde = make_product_sde(stg, seeded_unet(d, [256, 128, 64], 5), seeded_mnet(d, [128, 128], 6), gam, [256, 128, 64], [128, 128], DEV)
    x0 = torch.zeros(148 * 128 * 2, d, device=DEV)
    t1 = timeit(lambda: simulate.rollout(sde, x0, ts, 1.0, seed=1), 3)
    print(f"  K1 {t1:.2f} ms for {x0.shape[0]} paths x {K} steps", end="")
print()
