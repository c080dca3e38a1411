"""Sampling-loop parity on the GPU: Flow.denoise through the fused Euler / Heun kernels
against reference-generated trajectories (tests/golden/flow_misc.pt) and the oracle's Heun definition, CUDA-graph replay of
the loop against eager execution, and the classifier-free-guidance batching rule (SprintDiT must not batch)."""
import os

import pytest
import torch

from golden_util import load_fixture

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flow_misc.pt")


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


class ToyVelocity(torch.nn.Module):
    """Same field as oracle/make_golden_flow.py's ToyVelocity."""

    classifier_free = True

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(0.4))
        self.b = torch.nn.Parameter(torch.tensor(-0.3))

    def forward(self, x, timesteps, p=0.0, **_):
        s = torch.cos(3.0 * timesteps.float()).view(-1, 1, 1, 1)
        return {"x": self.a * x + self.b * s + (0.25 if p == 1 else 0.0)}


def test_euler_denoise_matches_reference_trajectories(cuda_device):
    import diffulab_b200 as dl

    fx = torch.load(GOLDEN, map_location="cpu", weights_only=False)
    model = ToyVelocity().cuda()
    for case in fx["denoise"]:
        flow = dl.Flow(n_steps=case["n"], sampling_method="euler")
        flow.set_steps(case["n"], shift=case["shift"])
        out = flow.denoise(model, {"x": case["x_init"].cuda()}, use_tqdm=False, guidance_scale=case["guidance"], return_intermediates=True)
        assert rel_l2(out["x"], case["x"]) < 2e-6
        assert rel_l2(out["xt"], case["xt"]) < 2e-6
        assert rel_l2(out["estimated_x0"], case["estimated_x0"]) < 2e-6


@pytest.mark.parametrize("guidance", [0.0, 3.0])
@pytest.mark.parametrize("shift", [None, 6.93])
def test_heun_matches_oracle(cuda_device, guidance, shift):
    """Heun = two reference-style velocity evaluations per step, fused predictor / corrector launches."""
    import diffulab_b200 as dl
    from oracle import dit_oracle as O

    assert dl.Flow.sampler_registry["heun"] is dl.Heun
    model = ToyVelocity().cuda()
    g = torch.Generator().manual_seed(3)
    x_init = torch.randn(4, 3, 8, 8, generator=g)
    flow = dl.Flow(n_steps=9, sampling_method="heun")
    flow.set_steps(9, shift=shift)
    out = flow.denoise(model, {"x": x_init.cuda()}, use_tqdm=False, guidance_scale=guidance, return_intermediates=True)

    def velocity(x, t, p):
        return 0.4 * x + (-0.3) * torch.cos(3.0 * torch.full((x.shape[0],), t)).view(-1, 1, 1, 1) + (0.25 if p == 1 else 0.0)

    ref = O.flow_denoise(velocity, x_init.clone(), 9, shift, guidance, method="heun")
    assert rel_l2(out["x"], ref) < 2e-6
    assert out["xt"].shape[1] == 10 and out["estimated_x0"].shape[1] == 9


def _small_dit(kind="mmdit"):
    import diffulab_b200 as dl

    torch.manual_seed(0)
    if kind == "sprint":
        m = dl.SprintDiT(simple_dit=True, input_channels=4, output_channels=4, inner_dim=128, embedding_dim=128, num_heads=2, mlp_ratio=4,
                         patch_size=2, encoder_depth=1, deep_layers_depth=2, decoder_depth=1, n_classes=10, classifier_free=True, drop_rate=0.5)
    else:
        m = dl.MMDiT(simple_dit=True, input_channels=4, output_channels=4, inner_dim=128, embedding_dim=128, num_heads=2, mlp_ratio=4,
                     patch_size=2, depth=3, n_classes=10, classifier_free=True)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m.cuda().eval()


@pytest.mark.parametrize("method", ["euler", "heun"])
@pytest.mark.parametrize("guidance", [0.0, 4.0])
def test_cuda_graph_loop_equals_eager(cuda_device, method, guidance):
    """The whole n-step loop replayed as one CUDA graph gives the eager result bit for bit, and a second replay with new
    inputs tracks the eager path again (static input buffers are refreshed)."""
    import diffulab_b200 as dl

    model = _small_dit()
    flow = dl.Flow(n_steps=6, sampling_method=method)
    flow.set_steps(6, shift=6.93)
    y = torch.tensor([1, 5, 7], device="cuda")
    for seed in (0, 1):
        x_init = torch.randn(3, 4, 16, 16, device="cuda", generator=torch.Generator("cuda").manual_seed(seed))
        flow.cuda_graph = False
        eager = flow.denoise(model, {"x": x_init.clone(), "y": y}, use_tqdm=False, guidance_scale=guidance)["x"]
        flow.cuda_graph = True
        graphed = flow.denoise(model, {"x": x_init.clone(), "y": y}, use_tqdm=False, guidance_scale=guidance)["x"]
        assert torch.equal(eager, graphed)
    assert len(flow._graphs) == 1


def test_sprint_cfg_is_not_batched_and_matches_reference(cuda_device):
    """SprintDiT's unconditional pass (p = 1) skips the deep layers (reference sprint.py:474-475); the batched [y; null]
    evaluation would run them. The reference trajectory pins the behaviour with batch_cfg on AND off."""
    import diffulab_b200 as dl
    from test_models_gpu import build_model, inputs_for

    fx = load_fixture("sprint_dit_cfg")
    e = fx["euler"]
    model = build_model(fx).eval()
    assert model.cfg_batchable is False
    outs = []
    for batch_cfg in (True, False):
        flow = dl.Flow(n_steps=e["n_steps"], sampling_method="euler", shift=e["shift"])
        flow.set_steps(e["n_steps"], shift=e["shift"])
        flow.batch_cfg = batch_cfg
        inp = {"x": e["x_init"].cuda(), **inputs_for(fx)}
        assert not flow._can_batch_cfg(model, inp)
        out = flow.denoise(model, inp, use_tqdm=False, guidance_scale=e["guidance"], return_intermediates=True)
        assert rel_l2(out["x"], e["x_final"]) < 5e-2 and rel_l2(out["xt"], e["xt"]) < 5e-2
        outs.append(out["x"])
    assert torch.equal(outs[0], outs[1])
    # a label-conditioned MMDiT does batch
    assert dl.Flow(n_steps=2)._can_batch_cfg(_small_dit(), {"x": e["x_init"].cuda(), "y": fx["y"].cuda()})


def test_batched_cfg_consumes_the_reference_rng_draw(cuda_device):
    """The reference's p = 1 pass draws torch.rand(labels.size()) on the device (nn.py:149); the batched path consumes the
    same draw, so the CUDA Philox stream seen by a stochastic sampler is identical with batch_cfg on and off."""
    import diffulab_b200 as dl

    model = _small_dit()
    y = torch.tensor([1, 5, 7], device="cuda")
    x = torch.randn(3, 4, 16, 16, device="cuda")
    after = []
    for batch_cfg in (True, False):
        flow = dl.Flow(n_steps=2)
        flow.batch_cfg = batch_cfg
        torch.manual_seed(123)
        with torch.inference_mode():
            flow._velocities(model, {"x": x, "y": y}, 0.5, 2.0)
        after.append(torch.rand(4, device="cuda"))
    assert torch.equal(after[0], after[1])
