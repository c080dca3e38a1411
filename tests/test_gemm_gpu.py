"""tcgen05 GEMM vs torch.matmul (fp32 reference of the same bf16 inputs). Tolerance: bf16 output rounding
(rel 2^-8) on top of fp32 accumulation-order noise -> relL2 <= 4e-3 (bf16 out) / 1e-5 (fp32 out)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


SHAPES = [
    # M, N, K
    (128, 128, 64),
    (128, 256, 128),
    (256, 192, 1152),
    (384, 3456, 1152),
    (512, 1152, 4608),
    (100, 72, 40),      # ragged everything (K multiple of 8 for the 16-byte TMA stride rule)
    (32, 16, 1152),     # skinny output (last layer)
    (4096, 1152, 16),   # tiny K (patch embed)
    (128, 6912, 1152),  # modulation
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tile_n", [0, 64, 128, 192, 256])
def test_gemm_nt_bf16(cuda_device, M, N, K, tile_n):
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b, tile_n=tile_n)
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16 and out.shape == (M, N)
    assert rel_l2(out, ref) < 4e-3
    out32 = ops.gemm(a, b, out_dtype=torch.float32, tile_n=tile_n)
    assert rel_l2(out32, ref) < 1e-5


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bias(cuda_device, M, N, K):
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ b.float().t() + bias
    out = ops.gemm(a, b, bias=bias, out_dtype=torch.float32)
    assert rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(256, 192, 128), (384, 1152, 3456), (1152, 1152, 4096), (1152, 16, 2048), (16, 1152, 512), (200, 136, 72)])
@pytest.mark.parametrize("tile_n", [0, 64, 128, 256])
def test_gemm_majors(cuda_device, a_mn, b_mn, M, N, K, tile_n):
    """dgrad (B read MN-major) and wgrad (both operands MN-major) straight from row-major tensors."""
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    a_in = a.t().contiguous() if a_mn else a
    b_in = b.t().contiguous() if b_mn else b
    out = ops.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, tile_n=tile_n)
    assert rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize("split_k", [1, 2, 3, 7, 64])
def test_gemm_accumulate_split_k(cuda_device, split_k):
    from diffulab_b200 import ops

    M, N, K = 1152, 320, 4096 + 64
    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.randn(K, M, device="cuda", generator=g).bfloat16()  # MN-major operands: wgrad layout
    b = torch.randn(K, N, device="cuda", generator=g).bfloat16()
    base = torch.randn(M, N, device="cuda", generator=g)
    out = base.clone()
    ops.gemm(a, b, a_mn=True, b_mn=True, out=out, accumulate=True, split_k=split_k)
    ref = base + a.float().t() @ b.float()
    assert rel_l2(out, ref) < 1e-5


def test_gemm_strided_views(cuda_device):
    """Operands and outputs that are column slices of packed buffers (qkv / modulation chunks)."""
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(11)
    packed = torch.randn(512, 3 * 384, device="cuda", generator=g).bfloat16()
    w = torch.randn(256, 384, device="cuda", generator=g).bfloat16()
    outbuf = torch.zeros(512, 2 * 256, device="cuda", dtype=torch.bfloat16)
    a = packed[:, 384:768]
    ops.gemm(a, w, out=outbuf[:, 256:])
    ref = a.float() @ w.float().t()
    assert rel_l2(outbuf[:, 256:], ref) < 4e-3
    assert outbuf[:, :256].abs().max().item() == 0.0


def test_gemm_large_perf_shape(cuda_device):
    """DiT-XL/2 qkv projection at per-GPU batch 32 (M = 8192): full persistent multi-wave path."""
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(13)
    a = torch.randn(8192, 1152, device="cuda", generator=g).bfloat16()
    b = (torch.randn(3456, 1152, device="cuda", generator=g) * 0.03).bfloat16()
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b)
    assert rel_l2(out, ref) < 4e-3


def test_gemm_bad_args(cuda_device):
    from diffulab_b200 import _lib, ops

    a = torch.zeros(16, 12, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.DlbError):
        ops.gemm(a, b)  # K=12 rows are 24 bytes: violates the 16-byte stride rule, must fail loudly


@pytest.mark.parametrize("M,F,K,with_bias", [(128, 128, 64, True), (100, 256, 128, False), (256, 128, 64, False), (300, 256, 192, True), (1024, 1152, 288, False), (4096, 4608, 1152, False)])
def test_gemm_swiglu_fused_matches_unfused(cuda_device, M, F, K, with_bias):
    """fc1 + SwiGLU epilogue == gemm followed by the standalone swiglu kernel, bit for bit (same bf16 rounding points)."""
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + F)
    a = (torch.randn(M, K, device="cuda", generator=g)).bfloat16()
    w = (torch.randn(2 * F, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(2 * F, device="cuda", generator=g) if with_bias else None
    h_ref = ops.gemm(a, w, bias=bias)
    act_ref = ops.swiglu_fwd(h_ref)
    h, act = ops.gemm_swiglu(a, w, bias)
    assert torch.equal(h, h_ref)
    assert torch.equal(act, act_ref)
    # and against an fp32 torch restatement
    hf = a.float() @ w.float().t() + (bias if bias is not None else 0)
    ref = torch.nn.functional.silu(hf[:, :F]) * hf[:, F:]
    assert ((act.float() - ref).norm() / ref.norm()).item() < 1e-2


@pytest.mark.parametrize("M,F,D", [(128, 128, 64), (100, 128, 192), (129, 256, 64), (300, 256, 192), (1000, 1152, 288), (4096, 4608, 1152)])
def test_gemm_swiglu_bwd_fused_matches_unfused(cuda_device, M, F, D):
    """fc2 dgrad + SwiGLU backward in the epilogue vs the two separate kernels (relL2: the unfused kernel uses the exact
    sigmoid, the epilogue tanh.approx) and vs an fp32 torch restatement."""
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + F + D)
    dy = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    w2 = (torch.randn(D, F, device="cuda", generator=g) * D ** -0.5).bfloat16()
    h = torch.randn(M, 2 * F, device="cuda", generator=g).bfloat16()
    ref_kernels = ops.swiglu_bwd(ops.gemm(dy, w2, b_mn=True), h)
    got = ops.gemm_swiglu_bwd(dy, w2, h)
    assert ((got.float() - ref_kernels.float()).norm() / ref_kernels.float().norm()).item() < 3e-3
    dact = (dy.float() @ w2.float())
    a, gg = h.float()[:, :F], h.float()[:, F:]
    sg = torch.sigmoid(a)
    ref = torch.cat([dact * gg * sg * (1 + a * (1 - sg)), dact * a * sg], 1)
    assert ((got.float() - ref).norm() / ref.norm()).item() < 1e-2


@pytest.mark.parametrize("tile_n", [128, 256])
@pytest.mark.parametrize("mode", ["fwd", "dgrad", "wgrad", "wgrad_splitk", "fwd_bias_f32"])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 384, 192), (1000, 520, 200), (4096, 1152, 1152)])
def test_gemm_cta_pair_matches_single_cta(cuda_device, M, N, K, mode, tile_n):
    """cta_group::2 kernel == single-CTA kernel bit for bit (same k-order of fp32 accumulation), all operand majors."""
    from diffulab_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    rnd = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()  # noqa: E731
    if mode in ("fwd", "fwd_bias_f32"):
        a, b, kw = rnd(M, K), rnd(N, K), {}
        if mode == "fwd_bias_f32":
            kw = dict(bias=torch.randn(N, device="cuda", generator=g), out_dtype=torch.float32)
    elif mode == "dgrad":
        a, b, kw = rnd(M, K), rnd(K, N), dict(b_mn=True)
    else:
        a, b, kw = rnd(K, M), rnd(K, N), dict(a_mn=True, b_mn=True)
    if mode.startswith("wgrad"):
        sk = 4 if mode == "wgrad_splitk" else 1
        ref = torch.zeros(M, N, device="cuda")
        got = torch.zeros(M, N, device="cuda")
        ops.gemm(a, b, out=ref, accumulate=True, split_k=1, tile_n=64, **kw)  # tile_n 64 exists only in the single-CTA kernel
        ops.gemm(a, b, out=got, accumulate=True, split_k=sk, tile_n=tile_n, pair=True, **kw)
        if sk == 1:
            assert torch.equal(got, ref)
        else:
            torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-3)
    else:
        ref = ops.gemm(a, b, tile_n=64, **kw)  # tile_n 64 exists only in the single-CTA kernel
        got = ops.gemm(a, b, tile_n=tile_n, pair=True, **kw)
        assert torch.equal(got, ref)
        if mode in ("fwd", "fwd_bias_f32"):  # K-major B: the 256 x 192 pair tile as well
            assert torch.equal(ops.gemm(a, b, tile_n=192, pair=True, **kw), ref)
