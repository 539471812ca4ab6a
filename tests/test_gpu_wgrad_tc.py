"""K3b (csrc/wgrad_tc.cu) on a synthetic scratch: every weight / bias gradient of the default UNet
equals sum_p dY[p] (x) Act[p] computed by torch in fp64 from the same fp32 operands (3xTF32 in the
kernel, fp32 accumulation over all points in TMEM: tolerance 1e-5 relative; a single TF32 pass gives ~3e-4)."""
import numpy as np
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"

FB = dict(XIN=0, R1=1, R2=9, R3=13, O2=15, O1=19, DY0=27, DO0=28, DY1=29, DO1=37, DY2=45, DO2=49, DZ3=53, DZ2=55,
          DZ1=59)
WIDTH = dict(XIN=32, R1=256, R2=128, R3=64, O2=128, O1=256, DY0=32, DO0=32, DY1=256, DO1=256, DY2=128, DO2=128,
             DZ3=64, DZ2=128, DZ1=256)
NFB = 67


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
    _lib.check(lib.socm_debug_wgrad_tc(scratch.data_ptr(), n_tiles, d, grad.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    grad = grad.cpu()
    dd = {k: v.double() for k, v in t.items()}
    x = dd["XIN"][:, :d + 1]
    pairs = [("DZ1", x), ("DZ2", dd["R1"]), ("DZ3", dd["R2"]), ("DO0", x), ("DO1", dd["R1"]), ("DO2", dd["R2"]),
             ("DY2", dd["R3"]), ("DY1", dd["O2"]), ("DY0", dd["O1"])]
    off = 0
    for l, (dy, act) in enumerate(pairs):
        dyv = dd[dy][:, :nout[l]]
        want_w = dyv.t() @ act
        want_b = dyv.sum(0)
        got_w = grad[off:off + nout[l] * nin[l]].reshape(nout[l], nin[l])
        off += nout[l] * nin[l]
        got_b = grad[off:off + nout[l]]
        off += nout[l]
        assert rel_l2(got_w, want_w) < 1e-5, (l, dy, rel_l2(got_w, want_w))
        assert rel_l2(got_b, want_b) < 1e-5, (l, dy, "bias", rel_l2(got_b, want_b))
