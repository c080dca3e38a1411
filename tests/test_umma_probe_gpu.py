"""Pins the non-swizzled UMMA shared-memory descriptor semantics (LBO = K-direction core-matrix stride, SBO =
MN-direction core-matrix stride, for K-major and MN-major operands) used by the tcgen05 attention kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 80), (80, 128), (256, 64), (16, 16)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_umma_noswizzle_descriptors(cuda_device, N, K, a_mn, b_mn):
    from diffulab_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(N + K)
    A = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = A.float() @ B.float().t()
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    errs = {}
    for swap in (0, 1):
        D = torch.zeros(128, N, device="cuda")
        rc = _lib.load_probes().dlb_umma_probe(Ain.data_ptr(), Bin.data_ptr(), D.data_ptr(), N, K, a_mn, b_mn, swap,
                                        torch.cuda.current_stream().cuda_stream)
        _lib.check_probe(rc, "dlb_umma_probe")
        torch.cuda.synchronize()
        errs[swap] = ((D - ref).norm() / ref.norm()).item()
    print("probe", N, K, a_mn, b_mn, errs)
    assert errs[0] < 1e-5, errs
