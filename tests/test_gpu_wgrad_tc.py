"""K3b (csrc/wgrad_tc.cu) on a synthetic scratch: every weight / bias gradient of the default UNet
equals sum_p dY[p] (x) Act[p] computed by torch in fp64 from the same fp32 operands (3xTF32 in the
kernel, fp32 accumulation over all points in TMEM: tolerance 1e-5 relative; a single TF32 pass gives ~3e-4)."""
import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"

FB = dict(XIN=0, R1=1, R2=9, R3=13, O2=15, Y1=19, DY0=27, DO0=28, DY1=29, DY2=37, DO2=41, DZ3=45, DZ2=47, DZ1=51)
WIDTH = dict(XIN=32, R1=256, R2=128, R3=64, O2=128, Y1=256, DY0=32, DO0=32, DY1=256, DY2=128, DO2=128,
             DZ3=64, DZ2=128, DZ1=256)
NFB = 59


def tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def build_scratch(tensors, n_tiles):
    """tensors[name]: (n_tiles*128, width) fp32 -> byte layout of loss_tc.cuh: feature block = 32 feature
    rows of 128 bytes (32 points), 16-byte chunks XOR-permuted by (feature & 7)."""
    words = torch.zeros(n_tiles, 4, NFB, 32, 32, dtype=torch.float32)       # tile, quarter, fb, feature, word
    f = torch.arange(32)
    pt = torch.arange(32)
    idx = (((pt[None, :] >> 2) ^ (f[:, None] & 7)) & 7) * 4 + (pt[None, :] & 3)   # (feature, point) -> word in the row
    for name, val in tensors.items():
        w = WIDTH[name]
        v = val.reshape(n_tiles, 4, 32, w // 32, 32).permute(0, 1, 3, 4, 2)       # tile, q, fb, feature, point
        blk = torch.zeros(n_tiles, 4, w // 32, 32, 32)
        blk.scatter_(4, idx[None, None, None].expand(n_tiles, 4, w // 32, 32, 32), v)
        words[:, :, FB[name]:FB[name] + w // 32] = blk
    return words.reshape(-1)


@pytest.mark.parametrize("d,n_tiles", [(10, 3), (1, 2), (20, 5), (10, 300)])
def test_wgrad_tc_matches_torch(d, n_tiles):
    from soc_matching_b200 import _lib
    lib = _lib.load()
    assert lib.socm_debug_wgrad_tile_bytes() == 4 * NFB * 4096
    g = torch.Generator().manual_seed(d * 100 + n_tiles)
    P = n_tiles * 128
    t = {k: torch.randn(P, w, generator=g) for k, w in WIDTH.items()}
    t["XIN"][:, d + 1:] = 0
    t["XIN"][:, 31] = 1.0
    t["DY0"][:, d:] = 0
    t["DO0"][:, d:] = 0
    scratch = build_scratch(t, n_tiles).to(DEV)
    nout = [256, 128, 64, d, 256, 128, 128, 256, d]
    nin = [d + 1, 256, 128, d + 1, 256, 128, 64, 128, 256]
    total = sum(o * i + o for o, i in zip(nout, nin))
    grad = torch.zeros(total, device=DEV)
    aux = torch.zeros(32 * 256 + 32, device=DEV)
    _lib.check(lib.socm_debug_wgrad_tc(scratch.data_ptr(), n_tiles, d, grad.data_ptr(), aux.data_ptr(),
                                       _lib.stream_ptr()))
    torch.cuda.synchronize()
    grad = grad.cpu()
    aux = aux.cpu()
    dd = {k: v.double() for k, v in t.items()}
    x = dd["XIN"][:, :d + 1]
    # res_1 (layer 4) is folded into up_0 (unet_tc.cuh): K3b leaves its gradient at zero and instead
    # accumulates S = d_y0^T r1 and sb = sum d_y0 in `aux`; up_0 (layer 8) receives d_y0^T y1 only.
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
        want_w = dyv.t() @ act
        want_b = dyv.sum(0)
        assert rel_l2(got_w, want_w) < 1e-5, (l, dy, rel_l2(got_w, want_w))
        assert rel_l2(got_b, want_b) < 1e-5, (l, dy, "bias", rel_l2(got_b, want_b))
    S = aux[:32 * 256].reshape(32, 256)
    want_S = dd["DY0"][:, :d].t() @ dd["R1"]
    assert rel_l2(S[:d], want_S) < 1e-5, rel_l2(S[:d], want_S)
    assert float(S[d:].abs().max()) == 0.0
    assert rel_l2(aux[32 * 256:32 * 256 + d], dd["DY0"][:, :d].sum(0)) < 1e-5
