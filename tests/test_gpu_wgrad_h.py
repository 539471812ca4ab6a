"""K3b of the fp16-split engine (csrc/wgrad_h.cu) on a synthetic scratch: every weight / bias gradient of the default
UNet equals sum_p dY[p] (x) Act[p] computed by torch in fp64 from the same fp32 operands.  The scratch holds fp16 hi / lo
planes of the operand-SCALED values (powers of two, one per tensor) in the MN-major core-matrix layout loss_h.cu writes;
the kernel forms hi hi + lo hi + hi lo on kind::f16 with fp32 accumulation and divides the scales out in the flush
(tolerance 1e-5 relative; hi hi alone gives ~3e-4)."""
import pytest
import torch

from helpers import rel_l2
from test_gpu_wgrad_tc import FB, NFB, WIDTH

pytestmark = pytest.mark.gpu
DEV = "cuda"
SCRATCH_T = ["XIN", "R1", "R2", "R3", "O2", "Y1", "DY0", "DO0", "DY1", "DY2", "DO2", "DZ3", "DZ2", "DZ1"]   # unet_h.cuh ScratchT


def build_scratch_h(tensors, scales, n_tiles):
    """tensors[name]: (n_tiles*128, width) fp32 -> fp16 scratch: tile, quarter, feature block (32 features), then
    [feature group of 8][plane hi|lo][point 0..31][feature in group] (loss_h.cu store_fb16)."""
    out = torch.zeros(n_tiles, 4, NFB, 4, 2, 32, 8, dtype=torch.float16)
    for name, val in tensors.items():
        w = WIDTH[name]
        v = val * scales[name]
        if name == "XIN":
            v = v.clone()
            v[:, 31] = 1.0   # the constant-1 feature is stored unscaled
        hi = v.to(torch.float16)
        lo = (v - hi.float()).to(torch.float16)
        for plane, x in enumerate((hi, lo)):
            x = x.reshape(n_tiles, 4, 32, w // 32, 4, 8).permute(0, 1, 3, 4, 2, 5)   # tile, q, fb, group, point, j
            out[:, :, FB[name]:FB[name] + w // 32, :, plane] = x
    return out.reshape(-1)


@pytest.mark.parametrize("d,n_tiles", [(10, 3), (1, 2), (15, 5), (10, 300)])
def test_wgrad_h_matches_torch(d, n_tiles):
    from soc_matching_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(d * 100 + n_tiles)
    P = n_tiles * 128
    t = {k: torch.randn(P, w, generator=g) for k, w in WIDTH.items()}
    t["XIN"][:, d + 1:] = 0
    t["DY0"][:, d:] = 0
    t["DO0"][:, d:] = 0
    exps = torch.randint(-2, 7, (len(SCRATCH_T),), generator=g)
    scales = {k: float(2.0 ** int(e)) for k, e in zip(SCRATCH_T, exps)}
    scales["DO0"] = scales["DY0"]   # d_o0 rides in the d_y0 units (loss_h.cu)
    scratch = build_scratch_h(t, scales, n_tiles).to(DEV)
    assert scratch.numel() * 2 == n_tiles * lib.socm_debug_wgrad_tile_bytes()
    sk = torch.tensor([scales[k] for k in SCRATCH_T], dtype=torch.float32, device=DEV)
    nout = [256, 128, 64, d, 256, 128, 128, 256, d]
    nin = [d + 1, 256, 128, d + 1, 256, 128, 64, 128, 256]
    total = sum(o * i + o for o, i in zip(nout, nin))
    grad = torch.zeros(total, device=DEV)
    aux = torch.zeros(32 * 256 + 32, device=DEV)
    _lib.check(lib.socm_debug_wgrad_h(scratch.data_ptr(), n_tiles, d, sk.data_ptr(), grad.data_ptr(), aux.data_ptr(),
                                      _lib.stream_ptr()))
    torch.cuda.synchronize()
    grad, aux = grad.cpu(), aux.cpu()
    dd = {k: v.double() for k, v in t.items()}
    x = dd["XIN"][:, :d + 1]
    # res_1 (layer 4) is folded into up_0: K3b leaves its gradient at zero and accumulates S = d_y0^T r1 and
    # sb = sum d_y0 in `aux`; up_0 (layer 8) receives d_y0^T y1 only (as wgrad_tc.cu)
    pairs = [("DZ1", x), ("DZ2", dd["R1"]), ("DZ3", dd["R2"]), ("DO0", x), None, ("DO2", dd["R2"]),
             ("DY2", dd["R3"]), ("DY1", dd["O2"]), ("DY0", dd["Y1"])]
    off = 0
    for l, pr in enumerate(pairs):
        got_w = grad[off:off + nout[l] * nin[l]].reshape(nout[l], nin[l])
        off += nout[l] * nin[l]
        got_b = grad[off:off + nout[l]]
        off += nout[l]
        if pr is None:
            assert float(got_w.abs().max()) == 0.0 and float(got_b.abs().max()) == 0.0
            continue
        dy, act = pr
        dyv = dd[dy][:, :nout[l]]
        assert rel_l2(got_w, dyv.t() @ act) < 1e-5, (l, dy, rel_l2(got_w, dyv.t() @ act))
        assert rel_l2(got_b, dyv.sum(0)) < 1e-5, (l, dy, "bias", rel_l2(got_b, dyv.sum(0)))
    S = aux[:32 * 256].reshape(32, 256)
    want_S = dd["DY0"][:, :d].t() @ dd["R1"]
    assert rel_l2(S[:d], want_S) < 1e-5, rel_l2(S[:d], want_S)
    assert float(S[d:].abs().max()) == 0.0
    assert rel_l2(aux[32 * 256:32 * 256 + d], dd["DY0"][:, :d].sum(0)) < 1e-5
