"""Live comparison of the oracle with the UNMODIFIED reference, where /root/reference is mounted (the build container;
skipped on the GPU box). Complements tests/test_oracle_golden.py (committed reference outputs): fresh weights and inputs,
fp32 AND bf16 autocast — the second pins the dtype handling of the oracle that the `ref_bf16` arm of the real-shape GPU parity
tests and bench.py's reference-GPU arm rely on (RMSNorm upcast, QKNorm cast to v, RoPE tables cast, cosine in fp32)."""
import pytest
import torch

from oracle.ref_shim import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not mounted (GPU box)")


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _setup(kind):
    import_reference()
    from diffulab.networks.denoisers.ddt import DDT
    from diffulab.networks.denoisers.mmdit import MMDiT
    from diffulab.networks.denoisers.sprint import SprintDiT
    from oracle.make_golden import rerandomize

    torch.manual_seed(3)
    if kind == "dit":
        kw = dict(simple_dit=True, input_channels=4, output_channels=4, inner_dim=96, embedding_dim=64, num_heads=2, mlp_ratio=4,
                  patch_size=2, depth=2, n_classes=10, classifier_free=True, rope_axes_dim=[16, 24])
        model = MMDiT(**kw)
    elif kind == "sprint":
        kw = dict(simple_dit=True, input_channels=4, output_channels=4, inner_dim=64, embedding_dim=64, num_heads=2, mlp_ratio=4,
                  patch_size=2, encoder_depth=1, deep_layers_depth=1, decoder_depth=1, n_classes=10, classifier_free=True, drop_rate=0.5)
        model = SprintDiT(**kw)
    else:
        kw = dict(simple_ddt=True, input_channels=4, output_channels=4, inner_dim=64, num_heads=2, mlp_ratio=4, patch_size=2,
                  encoder_depth=1, decoder_depth=1, n_classes=10, classifier_free=True)
        model = DDT(**kw)
    rerandomize(model, 5)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 4, 8, 8, generator=g)
    t = torch.rand(3, generator=g)
    y = torch.randint(0, 10, (3,), generator=g)
    return model.eval(), kw, x, t, y


@pytest.mark.parametrize("kind", ["dit", "sprint", "ddt"])
@pytest.mark.parametrize("autocast", [False, True])
def test_oracle_matches_reference_live(kind, autocast):
    from golden_util import oracle_forward
    from oracle import dit_oracle as O

    model, kw, x, t, y = _setup(kind)
    fx = {"kwargs": kw, "mm": False, "state_dict": model.state_dict(), "y": y, "context": None}
    O.set_round(None)
    O.set_fused_sdpa(True)
    try:
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            ref = model(x, t, y=y, p=0.0)["x"]
            got = oracle_forward(fx, dict(model.state_dict()), x, t, 0.0, {}, False)
    finally:
        O.set_fused_sdpa(False)
    assert got.dtype == ref.dtype
    # fp32: same arithmetic up to GEMM summation order; autocast: identical rounding points (bf16 ulp-level differences only)
    assert rel_l2(got, ref) < (5e-3 if autocast else 2e-5)  # bf16 eps = 7.8e-3
