"""Pins the swizzled head-slice tile layout of the tcgen05 attention kernels (csrc/attn_sw.cuh) on hardware: 3-D TMA boxes
(128-byte-swizzled main block + 32-byte-swizzled tail for head dim 72) consumed as K-major operands (Q K^T) and as MN-major
operands (P V), through the same device helpers the attention kernels call."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hd,H", [(72, 16), (64, 8), (32, 2), (128, 4), (40, 3)])
@pytest.mark.parametrize("packed", [1, 3])
def test_sw_tiles_k_major_and_mn_major(cuda_device, hd, H, packed):
    from diffulab_b200 import _lib

    lib = _lib.load_probes()
    g = torch.Generator(device="cuda").manual_seed(hd + H)
    rows = 512
    ld = packed * H * hd
    buf = torch.randn(rows, ld, device="cuda", generator=g).bfloat16()
    X = buf[:, (packed - 1) * H * hd:]  # e.g. the V third of a packed qkv projection (row pitch > H*hd)
    Y = torch.randn(rows, ld, device="cuda", generator=g).bfloat16()[:, : H * hd]  # same row pitch as X (the probe takes one ld)
    P = torch.randn(128, 64, device="cuda", generator=g).bfloat16()
    hdp = 64 if hd <= 64 else (80 if hd <= 80 else 128)
    for h, xrow, yrow in ((0, 0, 64), (H - 1, 256, 448), (H // 2, 384, 128)):
        D1 = torch.zeros(128, 64, device="cuda")
        D2 = torch.full((128, hdp), float("nan"), device="cuda")
        rc = lib.dlb_attn_sw_probe(X.data_ptr(), Y.data_ptr(), P.data_ptr(), D1.data_ptr(), D2.data_ptr(), rows, ld, H, hd, h, xrow, yrow, None,
                                   torch.cuda.current_stream().cuda_stream)
        _lib.check_probe(rc, "dlb_attn_sw_probe")
        torch.cuda.synchronize()
        xs = X[xrow:xrow + 128, h * hd:(h + 1) * hd].float()
        ys = Y[yrow:yrow + 64, h * hd:(h + 1) * hd].float()
        ref1 = xs @ ys.t()
        ref2 = P.float() @ ys
        e1 = ((D1 - ref1).norm() / ref1.norm()).item()
        e2 = ((D2[:, :hd] - ref2).norm() / ref2.norm()).item()
        assert e1 < 1e-5, (hd, h, "K-major", e1)
        assert e2 < 1e-5, (hd, h, "MN-major", e2)
        assert float(D2[:, hd:].abs().max()) == 0.0 if hd < hdp else True  # padded head-dim columns are exact zeros
