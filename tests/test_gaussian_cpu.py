"""Host logic of the Gaussian-diffusion drop-ins (no GPU): schedules, respacing, timestep_map and the float32
coefficient tables, against the reference's own outputs in tests/golden/gaussian.pt."""
import os

import pytest
import torch


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "gaussian.pt"), weights_only=False)


def test_schedules_and_respacing_match_reference(fx):
    from diffulab_b200 import GaussianDiffusion

    for tb in fx["tables"]:
        gd = GaussianDiffusion(**tb["kw"])
        if "set_steps" in tb:
            n, sched, sec = tb["set_steps"]
            gd.set_steps(n, schedule=sched, section_counts=sec)
            assert gd.steps == n
        assert gd.timestep_map == tb["timestep_map"]
        assert torch.equal(gd.betas, tb["betas"]) and gd.betas.dtype == tb["betas"].dtype
        assert torch.equal(gd.alphas_bar, tb["alphas_bar"])


def test_coefficient_table_rows_are_what_the_reference_extracts(fx):
    """Row t of the sampler table == the per-sample float32 scalars the oracle (pinned on the reference) derives."""
    from diffulab_b200 import GaussianDiffusion
    from oracle import gaussian_oracle as G

    gd = GaussianDiffusion(n_steps=1000, sampling_method="ddpm", sampler_parameters={"var_type": "fixed_large"})
    T = G.Tables(G.variance_schedule(1000))
    tab = gd.sampler._table
    t = torch.tensor([0, 1, 2, 500, 999])
    ex = lambda a: a[t].float()  # noqa: E731
    assert torch.equal(tab[t, 0], 1.0 / ex(T.sqrt_ab))
    assert torch.equal(tab[t, 1], (1 - ex(T.ab)).sqrt() / ex(T.sqrt_ab))
    assert torch.equal(tab[t, 4], ex(T.c1)) and torch.equal(tab[t, 5], ex(T.c2))
    seq = torch.cat([T.post_var[1:2], T.betas[1:]])
    assert torch.equal(tab[t, 6], ex(seq)) and torch.equal(tab[t, 7], torch.exp(0.5 * ex(torch.log(seq))))
    assert tab[0, 8] == 0 and tab[1, 8] == 1
    assert torch.equal(tab[t, 10], ex(T.ab_prev).sqrt())


def test_space_timesteps_and_errors():
    from diffulab_b200.diffuse import space_timesteps
    from diffulab_b200 import DDPM, GaussianDiffusion

    assert space_timesteps(1000, 10, ddim=True) == set(range(0, 1000, 100))  # the reference's own docstring example
    assert space_timesteps(300, "10,10,10") == space_timesteps(300, "10,10,10")
    assert len(space_timesteps(1000, 50)) == 50
    with pytest.raises(ValueError):
        space_timesteps(10, 11)
    with pytest.raises(ValueError):
        GaussianDiffusion(sampling_method="euler")
    with pytest.raises(ValueError):
        DDPM(mean_type="velocity")
    with pytest.raises(ValueError):
        DDPM(var_type="learned_sigma")
    t = GaussianDiffusion(n_steps=100).draw_timesteps(64)
    assert t.dtype == torch.int32 and t.shape == (64,) and int(t.min()) >= 0 and int(t.max()) < 100


def test_oracle_learned_variance_matches_reference():
    """the oracle's learned / learned_range DDPM step against the reference's own outputs (tests/golden/gaussian_learned.pt,
    made by oracle/make_golden_gaussian_learned.py): bit-exact on CPU"""
    from oracle import gaussian_oracle as G

    fl = torch.load(os.path.join(os.path.dirname(__file__), "golden", "gaussian_learned.pt"), weights_only=False)
    betas = G.variance_schedule(1000)
    assert len(fl["steps"]) == 12
    for c in fl["steps"]:
        out = G.ddpm_step(betas, c["mean_type"], c["var_type"], c["pred"], c["xt"], c["t"], c["noise"], c["clamp"])
        for k, v in c["out"].items():
            assert torch.equal(torch.nan_to_num(out[k]), torch.nan_to_num(v)), (c["var_type"], c["mean_type"], k)


def test_learned_range_table_columns():
    from diffulab_b200 import GaussianDiffusion
    from oracle import gaussian_oracle as G

    gd = GaussianDiffusion(n_steps=1000, sampler_parameters=dict(var_type="learned_range"))
    T = G.Tables(G.variance_schedule(1000))
    tab = gd.sampler._table
    assert torch.equal(tab[:, 14], T.post_logvar.float()) and torch.equal(tab[:, 15], T.betas.float().log())


def test_no_cpu_fallback():
    from diffulab_b200 import GaussianDiffusion

    gd = GaussianDiffusion(n_steps=10)
    with pytest.raises((ValueError, RuntimeError)):
        gd.add_noise(torch.zeros(2, 1, 4, 4), torch.tensor([1, 2], dtype=torch.int32), torch.zeros(2, 1, 4, 4))


def test_cfg_batching_is_only_used_for_label_conditioned_models():
    """Flow batches the conditional / unconditional evaluation into one forward only when the unconditional branch is
    'every label -> null class' (label-conditioned, classifier-free, no text context); host logic, no kernels involved."""
    from diffulab_b200 import Flow

    class M:
        label_embed = object()
        classifier_free = True
        n_classes = 10
        cfg_batchable = True

    class PathDrop(M):  # SprintDiT: p = 1 also skips the deep layers, so [y; null] batching would change the result
        cfg_batchable = False

    class NoCF(M):
        classifier_free = False

    class Ctx(M):
        label_embed = None

    f = Flow(n_steps=4)
    y = torch.zeros(2, dtype=torch.long)
    assert f._can_batch_cfg(M(), {"x": None, "y": y})
    assert not f._can_batch_cfg(M(), {"x": None})
    assert not f._can_batch_cfg(M(), {"x": None, "y": y, "initial_context": {"embeddings": None}})
    assert not f._can_batch_cfg(M(), {"x": None, "y": y, "x_context": torch.zeros(1)})
    assert not f._can_batch_cfg(NoCF(), {"x": None, "y": y})
    assert not f._can_batch_cfg(Ctx(), {"x": None, "y": y})
    assert not f._can_batch_cfg(PathDrop(), {"x": None, "y": y})
    import diffulab_b200 as dl

    assert dl.SprintDiT.cfg_batchable is False and dl.MMDiT.cfg_batchable is True and dl.DDT.cfg_batchable is True
    f.batch_cfg = False
    assert not f._can_batch_cfg(M(), {"x": None, "y": y})


def test_euler_maruyama_schedule_scalars():
    """sigma / std of the stochastic flow sampler are evaluated in Python floats exactly like the reference
    (samplers/flow/euler_meruyama.py:38-40); tmax = timesteps[1]."""
    from diffulab_b200 import Flow

    f = Flow(n_steps=10, sampling_method="euler_maruyama", sampler_parameters={"eta": 0.7})
    assert f.sampler.name == "euler_maruyama" and f.sampler.tmax == f.timesteps[1]
    t_curr, t_prev = f.timesteps[0], f.timesteps[1]
    sigma = ((t_curr / (1 - min(t_curr, f.sampler.tmax))) ** 0.5) * 0.7
    assert sigma > 0 and (t_curr - t_prev) > 0
    with pytest.raises((ValueError, RuntimeError)):
        f.sampler.step(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), t_curr, t_prev)  # CPU tensors: no fallback
