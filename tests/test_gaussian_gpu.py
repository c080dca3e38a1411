"""Gaussian-diffusion hot path on the GPU (a17 + DDPM / DDIM reverse steps) against the reference's own outputs
(tests/golden/gaussian.pt) and the pinned oracle. fp32 arithmetic in the reference's operation order:
tolerance 2e-6 relative (exp / log / sqrt / div ulps), checked element-wise."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "gaussian.pt"), weights_only=False)


def close(a, b, what, rtol=2e-6, atol=2e-6):
    a, b = torch.nan_to_num(a.float().cpu()), torch.nan_to_num(b.float().cpu().expand_as(a))
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")


class Toy(torch.nn.Module):
    classifier_free = True

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(0.3))
        self.b = torch.nn.Parameter(torch.tensor(-0.2))

    def forward(self, x, timesteps, p=0.0, **_):
        s = torch.sin(timesteps.float() / 100.0).view(-1, 1, 1, 1)
        return {"x": self.a * x + self.b * s * (0.5 if p == 1 else 1.0)}


def test_add_noise_matches_reference(cuda_device, fx):
    from diffulab_b200 import GaussianDiffusion

    gd = GaussianDiffusion(n_steps=1000)
    for c in fx["add_noise"]:
        xt, noise = gd.add_noise(c["x"].cuda(), c["t"].cuda(), c["noise"].cuda())
        close(xt, c["xt"], "add_noise")
        assert noise.data_ptr() != 0


def test_sampler_steps_match_reference(cuda_device, fx):
    from diffulab_b200 import GaussianDiffusion

    for c in fx["steps"]:
        gd = GaussianDiffusion(n_steps=1000, sampling_method=c["sampler"],
                               sampler_parameters=dict(mean_type=c["mean_type"], var_type=c["var_type"]))
        torch.manual_seed(c["seed"])
        torch.cuda.manual_seed(c["seed"])
        kw = {} if c["eta"] is None else {"eta": c["eta"]}
        # the CUDA generator draws different numbers than the CPU one: inject the recorded noise
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: c["noise"].to(t.device)
        try:
            out = gd.sampler.step(model_prediction=c["pred"].cuda(), timesteps=c["t"].cuda(), xt=c["xt"].cuda(), clamp_x=c["clamp"], **kw)
        finally:
            torch.randn_like = orig
        assert set(out) == set(c["out"]), (c["sampler"], c["eta"])
        tag = f"{c['sampler']}/{c['mean_type']}/{c['var_type']}/clamp={c['clamp']}/eta={c['eta']}"
        for k, v in c["out"].items():
            # log-probabilities divide by tiny variances at small t: compare relative to their magnitude
            close(out[k], v, f"{tag}:{k}", rtol=2e-5 if k == "logprob" else 2e-6, atol=1e-4 if k == "logprob" else 2e-6)


def test_learned_variance_steps_match_reference(cuda_device):
    """DDPM with var_type learned / learned_range (2C-channel prediction, per-element variance) against the reference's own
    step outputs; exp / log of a per-element value: 4e-6 relative (expf and logf are each <= 2 ulp)."""
    from diffulab_b200 import GaussianDiffusion

    fl = torch.load(os.path.join(os.path.dirname(__file__), "golden", "gaussian_learned.pt"), weights_only=False)
    for c in fl["steps"]:
        gd = GaussianDiffusion(n_steps=1000, sampling_method="ddpm", sampler_parameters=dict(mean_type=c["mean_type"], var_type=c["var_type"]))
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: c["noise"].to(t.device)
        try:
            out = gd.sampler.step(model_prediction=c["pred"].cuda(), timesteps=c["t"].cuda(), xt=c["xt"].cuda(), clamp_x=c["clamp"])
        finally:
            torch.randn_like = orig
        assert set(out) == set(c["out"])
        tag = f"ddpm/{c['mean_type']}/{c['var_type']}/clamp={c['clamp']}"
        for k, v in c["out"].items():
            close(out[k], v, f"{tag}:{k}", rtol=2e-5 if k == "logprob" else 4e-6, atol=1e-4 if k == "logprob" else 4e-6)
    # DDIM ignores the variance head (ddim.py:86: only x_start of _get_p_mean_var is used)
    c = fl["steps"][0]
    ddim = GaussianDiffusion(n_steps=1000, sampling_method="ddim", sampler_parameters=dict(mean_type=c["mean_type"], var_type="learned"))
    plain = GaussianDiffusion(n_steps=1000, sampling_method="ddim", sampler_parameters=dict(mean_type=c["mean_type"]))
    a = ddim.sampler.step(model_prediction=c["pred"].cuda(), timesteps=c["t"].cuda(), xt=c["xt"].cuda())
    b = plain.sampler.step(model_prediction=c["pred"][:, :3].contiguous().cuda(), timesteps=c["t"].cuda(), xt=c["xt"].cuda())
    assert torch.equal(a["x_prev"], b["x_prev"]) and torch.equal(a["estimated_x0"], b["estimated_x0"])
    with pytest.raises(ValueError):
        ddim.sampler.step(model_prediction=c["pred"][:, :3].contiguous().cuda(), timesteps=c["t"].cuda(), xt=c["xt"].cuda())


def test_sampler_accepts_bf16_predictions(cuda_device, fx):
    from diffulab_b200 import GaussianDiffusion
    from oracle import gaussian_oracle as G

    c = fx["steps"][0]
    gd = GaussianDiffusion(n_steps=1000)
    pred = c["pred"].cuda().bfloat16()
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: c["noise"].to(t.device)
    try:
        out = gd.sampler.step(model_prediction=pred, timesteps=c["t"].cuda(), xt=c["xt"].cuda())
    finally:
        torch.randn_like = orig
    ref = G.ddpm_step(G.variance_schedule(1000), "epsilon", "fixed_small", pred.float().cpu(), c["xt"], c["t"], c["noise"])
    close(out["x_prev"], ref["x_prev"], "bf16 pred x_prev")


def test_denoise_loops_match_reference(cuda_device, fx):
    """Whole reverse chains (respaced DDPM, DDIM with guidance, stochastic DDIM) around a toy denoiser: noise injected
    from a CPU generator seeded like the reference run, so every step must agree."""
    from diffulab_b200 import GaussianDiffusion

    for c in fx["denoise"]:
        gd = GaussianDiffusion(n_steps=1000 if c["sampler"] == "ddpm" else c["n"], sampling_method=c["sampler"])
        if c["sampler"] == "ddpm":
            gd.set_steps(c["n"])
        assert gd.timestep_map == c["timestep_map"]
        model = Toy().cuda()
        torch.manual_seed(c["seed"])
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: torch.randn(t.shape).to(t.device)  # CPU Philox stream, as in the fixture
        try:
            with torch.no_grad():
                out = gd.denoise(model, {"x": c["x_init"].cuda()}, use_tqdm=False, clamp_x=True, guidance_scale=c["guidance"],
                                 sampler_args=c["sampler_args"], return_intermediates=True)
        finally:
            torch.randn_like = orig
        assert set(out) == set(c["out"]), (c["sampler"], set(out), set(c["out"]))
        for k in ("x", "xt", "estimated_x0"):
            close(out[k], c["out"][k], f"denoise {c['sampler']} {k}", rtol=1e-4, atol=1e-4)


def test_compute_loss_and_grads_match_reference(cuda_device, fx):
    from diffulab_b200 import GaussianDiffusion

    c = fx["loss"]
    gd = GaussianDiffusion(n_steps=1000)
    model = Toy().cuda()
    loss = gd.compute_loss(model, {"x": c["x"].cuda()}, c["t"].cuda(), c["noise"].cuda())["loss"]
    loss.backward()
    close(loss, c["loss"], "loss", rtol=1e-5)
    close(model.a.grad, c["grad_a"], "grad a", rtol=1e-4)
    close(model.b.grad, c["grad_b"], "grad b", rtol=1e-4)


def test_diffuser_registry_builds_gaussian(cuda_device):
    from diffulab_b200 import Diffuser, GaussianDiffusion

    d = Diffuser(Toy().cuda(), sampling_method="ddim", model_type="gaussian_diffusion", n_steps=25)
    assert isinstance(d.diffusion, GaussianDiffusion) and d.diffusion.sampler.name == "ddim"


def test_euler_maruyama_steps_match_reference(cuda_device, fx):
    from diffulab_b200 import Flow

    for c in fx["em_steps"]:
        fl = Flow(n_steps=c["n"], sampling_method="euler_maruyama", sampler_parameters={"eta": c["eta"]})
        assert fl.timesteps[c["idx"]] == c["t_curr"] and fl.timesteps[c["idx"] + 1] == c["t_prev"]
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: c["noise"].to(t.device)
        try:
            given = None if c["x_prev_in"] is None else c["x_prev_in"].cuda()
            out = fl.sampler.step(c["x_t"].cuda(), c["v"].cuda(), c["t_curr"], c["t_prev"], x_prev=given)
        finally:
            torch.randn_like = orig
        assert set(out) == set(c["out"])
        for k, v in c["out"].items():
            close(out[k], v, f"euler_maruyama eta={c['eta']} idx={c['idx']}:{k}", rtol=2e-5 if k == "logprob" else 2e-6, atol=2e-5 if k == "logprob" else 2e-6)


def test_euler_maruyama_denoise_matches_reference(cuda_device, fx):
    from diffulab_b200 import Flow

    c = fx["em_denoise"]

    class ToyFlow(Toy):
        def forward(self, x, timesteps, p=0.0, **_):
            s = torch.sin(timesteps.float() * 3.0).view(-1, 1, 1, 1)
            return {"x": self.a * x + self.b * s * (0.5 if p == 1 else 1.0)}

    fl = Flow(n_steps=c["n"], sampling_method="euler_maruyama", sampler_parameters={"eta": c["eta"]})
    torch.manual_seed(c["seed"])
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: torch.randn(t.shape).to(t.device)
    try:
        out = fl.denoise(ToyFlow().cuda(), {"x": c["x_init"].cuda()}, use_tqdm=False, guidance_scale=c["guidance"], return_intermediates=True)
    finally:
        torch.randn_like = orig
    assert set(out) == set(c["out"])
    for k, v in c["out"].items():
        assert tuple(out[k].shape) == tuple(v.shape), (k, out[k].shape, v.shape)
        close(out[k], v, f"euler_maruyama denoise {k}", rtol=1e-4, atol=1e-4)
