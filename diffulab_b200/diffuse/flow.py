"""Rectified-flow formalisation: drop-in for reference diffuse/modelizations/flow.py:16-524 (same constructor,
`draw_timesteps`, `add_noise`, `compute_loss`, `get_v`, `one_step_denoise`, `denoise`, `set_steps`).
Device math runs in fused kernels: x_t = (1-t) x0 + t eps (one kernel); loss = mean((eps - x0 - pred)^2) and its
gradient (one kernel each, no materialised target); CFG combine + Euler update (one kernel)."""

from __future__ import annotations

from typing import Any, Literal, cast

import torch
from torch import Tensor

from .. import ops
from ..denoisers.common import Denoiser, ModelInput, ModelOutput
from ..losses.common import LossFunction
from .diffusion import Diffusion, SamplingOutput
from .samplers import Euler, EulerMaruyama, Heun, StepResult


class _FlowLossFn(torch.autograd.Function):
    """mean_b(mean_chw(((eps - x0) - v)^2)), v = pred (v-prediction) or (x_t - pred) / t (x-prediction)."""

    @staticmethod
    def forward(ctx, pred: Tensor, x0: Tensor, eps: Tensor, xt: Tensor | None, t: Tensor | None):
        pred = pred.contiguous()
        ctx.save_for_backward(pred, x0, eps, xt, t)
        return ops.mse_fwd(pred, x0, eps, xt, t)

    @staticmethod
    def backward(ctx, gout: Tensor):
        pred, x0, eps, xt, t = ctx.saved_tensors
        return ops.mse_bwd(pred, x0, eps, gout.to(torch.float32).contiguous(), xt, t), None, None, None, None


class Flow(Diffusion):
    sampler_registry = {"euler": Euler, "euler_maruyama": EulerMaruyama, "heun": Heun}

    def __init__(
        self,
        n_steps: int = 50,
        sampling_method: Literal["euler", "euler_maruyama", "heun"] = "euler",
        schedule: Literal["linear"] = "linear",
        latent_diffusion: bool = False,
        logits_normal: bool = False,
        shift: float | None = None,
        sampler_parameters: dict[str, Any] = {},
        prediction_type: Literal["v", "x"] = "v",
    ) -> None:
        assert prediction_type in ["v", "x"], "prediction_type must be either 'v' or 'x', noise prediction not supported yet for flow models"
        self.shift = shift
        super().__init__(n_steps=n_steps, sampling_method=sampling_method, schedule=schedule, latent_diffusion=latent_diffusion,
                         sampler_parameters=sampler_parameters)
        self.logits_normal = logits_normal
        self.shift = shift
        self.x_prediction = prediction_type == "x"

    @staticmethod
    def _shift_timestep(t: Tensor | float, alpha: float) -> Tensor | float:
        return alpha * t / (1 + (alpha - 1) * t)

    def set_steps(self, n_steps: int, schedule: str = "linear", shift: float | None = None) -> None:
        """reference flow.py:101-135 (note: like the reference, this overwrites self.shift with the argument)"""
        self.shift = shift
        if schedule != "linear":
            raise NotImplementedError("Only linear schedule is supported for the moment")
        self.schedule = schedule
        timesteps: list[float] = torch.linspace(1, 0, n_steps + 1).tolist()
        if self.shift is not None:
            timesteps = [self._shift_timestep(t, self.shift) for t in timesteps]  # type: ignore
        self.timesteps = timesteps
        self.steps = n_steps
        self.sampler.set_steps(self.timesteps)

    def at(self, timesteps: Tensor) -> Tensor:
        return torch.ones_like(timesteps) - timesteps

    def bt(self, timesteps: Tensor) -> Tensor:
        return timesteps

    def draw_timesteps(self, batch_size: int) -> Tensor:
        """reference flow.py:168-197: drawn on the CPU generator exactly as the reference does."""
        if self.logits_normal:
            t = torch.sigmoid(torch.randn((batch_size), dtype=torch.float32))
        else:
            t = torch.rand((batch_size), dtype=torch.float32)
        if self.shift is not None:
            t = self._shift_timestep(t, self.shift)  # type: ignore
        if self.x_prediction:
            t = t.clamp(min=0.05)
        return t

    def add_noise(self, x: Tensor, timesteps: Tensor, noise: Tensor | None = None) -> tuple[Tensor, Tensor]:
        if noise is None:
            noise = torch.randn_like(x)
        assert noise.shape == x.shape
        assert timesteps.shape[0] == x.shape[0]
        t = timesteps.to(device=x.device, dtype=torch.float32).contiguous()
        z_t = ops.interp(x.float().contiguous(), noise.float().contiguous(), self.at(t).contiguous(), self.bt(t).contiguous())
        return z_t, noise

    def get_v(self, model: Denoiser, model_inputs: ModelInput, t_curr: float) -> Tensor:
        device = next(model.parameters()).device
        dtype = next(model.parameters()).dtype
        timesteps = torch.full((model_inputs["x"].shape[0],), t_curr, device=device, dtype=dtype)
        prediction = model(**model_inputs, timesteps=timesteps)["x"]
        if self.x_prediction:
            return (model_inputs["x"] - prediction) / max(t_curr, 0.05)
        return prediction

    batch_cfg: bool = True  # run the two classifier-free-guidance evaluations as one batched forward where possible

    def _can_batch_cfg(self, model: Denoiser, model_inputs: ModelInput) -> bool:
        """Batching [x; x] with labels [y; null] equals the reference's two calls (p = 0, then p = 1) only when p touches
        nothing but the label. SprintDiT's p = 1 pass also replaces the deep path by mask tokens (path-drop guidance,
        reference sprint.py:474-475), so models advertise `cfg_batchable` and SprintDiT opts out."""
        return (self.batch_cfg and getattr(model, "cfg_batchable", False) and model_inputs.get("y") is not None
                and model_inputs.get("initial_context") is None and model_inputs.get("x_context") is None
                and getattr(model, "label_embed", None) is not None and getattr(model, "classifier_free", False)
                and getattr(model, "n_classes", None) is not None)

    def _velocities(self, model: Denoiser, model_inputs: ModelInput, t: float, guidance_scale: float) -> tuple[Tensor, Tensor | None]:
        """(v, v_dropped) at time t as the reference evaluates them (flow.py:254-259): p = 0, and p = 1 when guiding."""
        if guidance_scale > 0 and self._can_batch_cfg(model, model_inputs):
            # Label-conditioned classifier-free guidance: the conditional and the unconditional evaluation run as ONE
            # forward over [x; x] with labels [y; null]. Samples do not interact inside the denoiser (tests: batch-slice
            # independence), so the halves equal the two separate calls; the label dropout at p = 1 is deterministic.
            x = model_inputs["x"]
            y = model_inputs["y"]
            # the reference's p = 1 pass draws torch.rand(labels.size()) (nn.py:149); consume it so that seeded stochastic
            # samplers see the same Philox stream as the unbatched path
            torch.rand(y.size(), device=y.device)
            both = {**model_inputs, "x": torch.cat([x, x], 0), "y": torch.cat([y, torch.full_like(y, model.n_classes)], 0), "p": 0}
            v, v_dropped = self.get_v(model, ModelInput(both), t).chunk(2, 0)
            return v, v_dropped
        v = self.get_v(model, ModelInput({**model_inputs, "p": 0}), t)
        v_dropped = self.get_v(model, {**model_inputs, "p": 1}, t) if guidance_scale > 0 else None
        return v, v_dropped

    def one_step_denoise(self, model: Denoiser, model_inputs: ModelInput, t_prev: float, t_curr: float, guidance_scale: float,
                         sampler_args: dict[str, Any] = {}) -> StepResult:
        x = model_inputs["x"]
        v, v_dropped = self._velocities(model, model_inputs, t_curr, guidance_scale)
        if isinstance(self.sampler, Heun) and not sampler_args:
            if not self.sampler.needs_corrector(t_prev):
                return self.sampler.step_cfg(x, v, v_dropped, guidance_scale, t_curr, t_prev)
            # predictor (CFG combine + Euler update in one launch, also yields the combined v1), second evaluation at
            # (x_pred, t_prev), corrector x - (v1 + v2)/2 dt in one launch
            x_pred, v1 = self.sampler.predict_cfg(x, v, v_dropped, guidance_scale, t_curr, t_prev)
            v2, v2_dropped = self._velocities(model, ModelInput({**model_inputs, "x": x_pred}), t_prev, guidance_scale)
            if v2_dropped is not None:
                v2 = self.sampler.combine(x_pred, v2, v2_dropped, guidance_scale)
            return self.sampler.correct(x, v1, v2, t_curr, t_prev)
        if v_dropped is not None:
            if isinstance(self.sampler, Euler) and not sampler_args:
                # CFG combine v_u + g (v - v_u) and the Euler update in one kernel (flow.py:259 + euler.py:37-39)
                return self.sampler.step_cfg(x, v, v_dropped, guidance_scale, t_curr, t_prev)
            v = v_dropped + guidance_scale * (v - v_dropped)
        return self.sampler.step(x, v, t_curr, t_prev, **sampler_args)

    def compute_loss(self, model: Denoiser, model_inputs: ModelInput, timesteps: Tensor, noise: Tensor | None = None,
                     extra_losses: list[LossFunction] = [], extra_args: dict[str, Any] = {}) -> dict[str, Tensor]:
        x_0 = model_inputs["x"].float().contiguous()
        model_inputs["x"], noise = self.add_noise(x_0, timesteps, noise)  # mutates the caller's dict like the reference
        t_dev = timesteps.to(device=x_0.device, dtype=torch.float32).contiguous()
        prediction: ModelOutput = model(**model_inputs, timesteps=t_dev)
        noise = noise.float().contiguous()
        if self.x_prediction:
            loss = _FlowLossFn.apply(prediction["x"], x_0, noise, model_inputs["x"], t_dev)
        else:
            loss = _FlowLossFn.apply(prediction["x"], x_0, noise, None, None)
        loss_dict = {"loss": loss}
        for extra_loss in extra_losses:
            loss_dict[extra_loss.name] = cast(Tensor, extra_loss(**extra_args))
        return loss_dict

    # Replay the whole sampling loop as ONE CUDA graph (Euler / Heun, no intermediates): the per-step time values are plain
    # kernel arguments, so the n-step loop (n or 2n denoiser evaluations + fused updates) is captured once per
    # (model, shapes, schedule, guidance) and replayed; at small batch the loop is launch-bound and this removes the host.
    cuda_graph: bool = False

    def _graph_key(self, model: Denoiser, model_inputs: ModelInput, guidance_scale: float) -> tuple:
        def sig(v: Any) -> Any:
            if isinstance(v, Tensor):
                return (tuple(v.shape), str(v.dtype), str(v.device))
            if isinstance(v, dict):
                return tuple((k, sig(x)) for k, x in sorted(v.items()))
            return v
        from .. import blocks as K

        # captured kernels read the bf16 weight shadows by address: any parameter update (version bump) or shadow rebuild
        # must invalidate the graph
        wver = (sum(p._version for p in model.parameters()), K._shadow_epoch)
        return (id(model), wver, model.training, type(self.sampler).__name__, tuple(self.timesteps), float(guidance_scale), self.x_prediction,
                self.batch_cfg, sig(dict(model_inputs)))

    def _denoise_graphed(self, model: Denoiser, model_inputs: ModelInput, guidance_scale: float) -> Tensor:
        def clone(v: Any) -> Any:
            if isinstance(v, Tensor):
                return v.clone()
            if isinstance(v, dict):
                return {k: clone(x) for k, x in v.items()}
            return v

        def copy_into(dst: Any, src: Any) -> None:
            if isinstance(dst, Tensor):
                dst.copy_(src)
            elif isinstance(dst, dict):
                for k in dst:
                    copy_into(dst[k], src[k])

        cache = self.__dict__.setdefault("_graphs", {})
        key = self._graph_key(model, model_inputs, guidance_scale)
        ent = cache.get(key)
        if ent is None:
            static_in = clone(dict(model_inputs))
            ts = list(zip(self.timesteps[:-1], self.timesteps[1:]))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture: weight shadows, RoPE tables, tensor maps, allocator
                self.one_step_denoise(model, ModelInput(dict(static_in)), t_curr=ts[0][0], t_prev=ts[0][1], guidance_scale=guidance_scale)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                inp = dict(static_in)
                for t_curr, t_prev in ts:
                    inp["x"] = self.one_step_denoise(model, ModelInput(inp), t_curr=t_curr, t_prev=t_prev, guidance_scale=guidance_scale)["x_prev"]
                static_out = inp["x"]
            if len(cache) >= 8:
                cache.clear()
            ent = cache[key] = (graph, static_in, static_out)
        graph, static_in, static_out = ent
        copy_into(static_in, dict(model_inputs))
        graph.replay()
        return static_out.clone()

    @torch.inference_mode()
    def denoise(self, model: Denoiser, model_inputs: ModelInput, data_shape: tuple[int, ...] | None = None, use_tqdm: bool = True,
                clamp_x: bool = False, guidance_scale: float = 0, sampler_args: dict[str, Any] = {},
                return_intermediates: bool = False) -> SamplingOutput:
        device = next(model.parameters()).device
        dtype = next(model.parameters()).dtype
        if "x" not in model_inputs:
            assert data_shape is not None, "'data_shape' must be provided if 'x' is not in model_inputs"
            model_inputs["x"] = torch.randn(data_shape, device=device, dtype=dtype)
        all_x0: list[Tensor] = []
        all_xt: list[Tensor] = [model_inputs["x"]]
        all_xt_mean: list[Tensor] = []
        all_xt_std: list[Tensor] = []
        all_logprobs: list[Tensor] = []
        if self.cuda_graph and not return_intermediates and not sampler_args and isinstance(self.sampler, (Euler, Heun)):
            model_inputs["x"] = self._denoise_graphed(model, model_inputs, guidance_scale)
            if clamp_x:
                model_inputs["x"] = model_inputs["x"].clamp(-1, 1)
            return {"x": model_inputs["x"]}
        for t_curr, t_prev in zip(self.timesteps[:-1], self.timesteps[1:]):
            step_output = self.one_step_denoise(model, model_inputs, t_curr=t_curr, t_prev=t_prev, guidance_scale=guidance_scale,
                                                sampler_args=sampler_args)
            model_inputs["x"] = step_output["x_prev"]
            if return_intermediates:
                all_xt.append(step_output["x_prev"])
                all_x0.append(step_output["estimated_x0"])
                if "x_prev_mean" in step_output:
                    all_xt_mean.append(step_output["x_prev_mean"])
                if "x_prev_std" in step_output:
                    all_xt_std.append(step_output["x_prev_std"])
                if "logprob" in step_output:
                    all_logprobs.append(step_output["logprob"])
        if clamp_x:
            model_inputs["x"] = model_inputs["x"].clamp(-1, 1)
        out: SamplingOutput = {"x": model_inputs["x"]}
        if return_intermediates:
            out["xt"] = torch.stack(all_xt, dim=1)
            out["estimated_x0"] = torch.stack(all_x0, dim=1)
            if all_xt_mean:
                out["xt_mean"] = torch.stack(all_xt_mean, dim=1)
            if all_xt_std:
                out["xt_std"] = torch.stack(all_xt_std, dim=0)  # dim 0, as in the reference (scalar per step)
            if all_logprobs:
                out["logprob"] = torch.stack(all_logprobs, dim=1)
        return out
