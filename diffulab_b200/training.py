"""Training-side runtime of the hot path: flat parameter / gradient / bf16-shadow storage, a fused AdamW step,
the bucketed data-parallel gradient all-reduce overlapped with backward, and `training_step`.

Reference: BaseTrainer.training_step training/trainers/base_trainer.py:104-153 (zero_grad -> draw t -> set p ->
compute_loss -> backward -> optimizer.step) executed under HF Accelerate's DDP (training/trainers/common.py:103-109:
fp32 gradients, bucketed all-reduce overlapped with backward, mean over ranks). Here: one process per GPU
(torchrun), gradients are written by the wgrad kernels straight into ONE flat fp32 buffer whose contiguous
buckets are all-reduced (NCCL over NVLink/NVSwitch) as soon as every gradient of a bucket has been produced.
"""

from __future__ import annotations

from typing import Any, Iterable

import torch
import torch.distributed as dist
from torch import Tensor, nn

from . import blocks as K
from . import ops

_ALIGN = 64  # elements; keeps every parameter slice 256-byte (fp32) / 128-byte (bf16) aligned for TMA


class FlatParams:
    """Re-homes parameters into one flat fp32 buffer (+ flat fp32 grads, + flat bf16 shadow used by the GEMMs).
    Parameters keep their identity, names and shapes (they become views), so state_dict / EMA / hooks still work."""

    def __init__(self, params: Iterable[nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        assert dev.type == "cuda", "FlatParams needs CUDA parameters (diffulab_b200 has no CPU path)"
        self.offsets: list[int] = []
        off = 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.flat_p = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_shadow = torch.empty(off, device=dev, dtype=torch.bfloat16)
        self._versions: list[int] = []
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                n = p.numel()
                self.flat_p[o : o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o : o + n].view(p.shape)
                p.grad = self.flat_g[o : o + n].view(p.shape)
        self.refresh_shadows()
        for p, o in zip(self.params, self.offsets):
            cols = p.numel() // p.shape[0] if p.dim() > 1 else p.numel()
            if p.dim() > 1 and cols % 8 == 0:
                K.install_shadow(p, self.flat_shadow[o : o + p.numel()].view(p.shape[0], cols))

    def refresh_shadows(self) -> None:
        ops._lib_call("dlb_cast_f32_bf16", self.flat_p.data_ptr(), self.flat_shadow.data_ptr(), 1, self.total, self.total, ops._stream())
        self._versions = [p._version for p in self.params]
        K.bump_shadow_epoch()

    def check_versions(self) -> None:
        """Parameters modified through the torch API (load_state_dict, copy_) since the last sync -> resync shadows."""
        if any(p._version != v for p, v in zip(self.params, self._versions)):
            self.refresh_shadows()

    def zero_grad(self) -> None:
        self.flat_g.zero_()
        for p, o in zip(self.params, self.offsets):  # re-attach if someone set .grad to None
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o : o + p.numel()].view(p.shape)


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction; no amsgrad) as ONE kernel over the flat
    parameter buffer; the same pass writes the bf16 shadow consumed by the next forward and, optionally, an EMA
    copy (reference: torch.optim.AdamW built from configs/optimizer/adamw.yaml + ema_pytorch, base_trainer.py:149-153).
    `state_dict()` keeps torch's per-parameter layout (step / exp_avg / exp_avg_sq)."""

    def __init__(self, params, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, grad_scale: float = 1.0):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_scale = grad_scale
        self.stores: list[FlatParams] = []
        self._steps: list[int] = []
        self._m: list[Tensor] = []
        self._v: list[Tensor] = []
        for group in self.param_groups:
            st = FlatParams(group["params"])
            self.stores.append(st)
            self._steps.append(0)
            m, v = torch.zeros_like(st.flat_p), torch.zeros_like(st.flat_p)
            step_t = torch.tensor(0.0)  # one shared counter tensor per group (torch keeps one per parameter)
            self._step_t = getattr(self, "_step_t", []) + [step_t]
            self._m.append(m)
            self._v.append(v)
            for p, o in zip(st.params, st.offsets):
                n = p.numel()
                self.state[p] = {"step": step_t, "exp_avg": m[o : o + n].view(p.shape), "exp_avg_sq": v[o : o + n].view(p.shape)}

    def zero_grad(self, set_to_none: bool = True) -> None:  # noqa: ARG002 - gradients live in the flat buffer
        for st in self.stores:
            st.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, (group, st) in enumerate(zip(self.param_groups, self.stores)):
            st.check_versions()
            self._steps[gi] += 1
            b1, b2 = group["betas"]
            ops.adamw_step(st.flat_p, st.flat_g, self._m[gi], self._v[gi], st.flat_shadow, lr=float(group["lr"]), beta1=b1, beta2=b2,
                           eps=group["eps"], weight_decay=group["weight_decay"], step=self._steps[gi], grad_scale=self.grad_scale)
            self._step_t[gi] += 1
        K.bump_shadow_epoch()  # version-cached (padded) shadows must be rebuilt; installed flat shadows stay valid
        return loss


class GradReducer:
    """Bucketed gradient all-reduce (mean over ranks) launched during backward.

    Buckets are contiguous ranges of a FlatParams gradient buffer (or per-parameter tensors when no flat store is
    given). `blocks.set_grad_ready_hook` reports each parameter the moment its gradient kernels have been enqueued;
    when all parameters of a bucket are ready the bucket's all-reduce is issued with async_op=True (c10d orders it
    after the compute stream's current position and runs it on the communicator's own stream), so communication
    overlaps the rest of backward. Parameters that never receive a gradient (reference quirk) are
    tolerated: their buckets are flushed, in a fixed order, by `finish()`."""

    def __init__(self, stores: list[FlatParams] | None = None, params: list[nn.Parameter] | None = None,
                 bucket_mb: float = 128.0, process_group: Any = None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets: list[Tensor] = []
        self.bucket_of: dict[int, int] = {}
        self.pending_init: list[int] = []
        cap = int(bucket_mb * 1024 * 1024 / 4)
        if stores:
            for st in stores:
                start, count, members = None, 0, []
                # walk parameters in REVERSE registration order: backward produces gradients roughly last-to-first
                for p, o in reversed(list(zip(st.params, st.offsets))):
                    n = (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
                    if start is None:
                        end = o + n
                    start = o
                    count += n
                    members.append(p)
                    if count >= cap:
                        self._add_bucket(st.flat_g[start:end], members)
                        start, count, members = None, 0, []
                if members:
                    self._add_bucket(st.flat_g[start:end], members)
        else:
            assert params is not None
            for p in reversed([q for q in params if q.requires_grad]):
                self._add_bucket(K.gbuf(p).view(-1), [p])
        self.pending = list(self.pending_init)
        self.launched = [False] * len(self.buckets)
        self.works: list[Any] = []
        self.active = False

    def _add_bucket(self, tensor: Tensor, members: list[nn.Parameter]) -> None:
        bi = len(self.buckets)
        self.buckets.append(tensor)
        for p in members:
            self.bucket_of[id(p)] = bi
        self.pending_init.append(len(members))

    def begin(self) -> None:
        self.pending = list(self.pending_init)
        self.launched = [False] * len(self.buckets)
        self.works = []
        self.active = True
        K.set_grad_ready_hook(self._ready)

    def _launch(self, bi: int) -> None:
        if self.launched[bi]:
            return
        self.launched[bi] = True
        if self.world == 1:
            return
        t = self.buckets[bi]
        if t.is_cuda:
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
        else:  # gloo (CPU tests of the host logic) has no AVG
            w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self.works.append((w, t))

    def _ready(self, p: Tensor) -> None:
        if not self.active:
            return
        bi = self.bucket_of.get(id(p))
        if bi is None:
            return
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def finish(self) -> None:
        """Flush buckets holding never-produced gradients (fixed order on every rank), then wait for all of them."""
        K.set_grad_ready_hook(None)
        self.active = False
        for bi in range(len(self.buckets)):
            self._launch(bi)
        for w in self.works:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()
        self.works = []


def training_step(diffuser, optimizer: torch.optim.Optimizer, batch: dict[str, Any], p_classifier_free_guidance: float = 0.0,
                  reducer: GradReducer | None = None, scheduler: Any | None = None) -> dict[str, Tensor]:
    """One optimisation step with the reference's order of operations (base_trainer.py:138-153). Returns the loss
    dict as DEVICE tensors: the reference's per-step `loss.item()` host sync is left to the caller."""
    optimizer.zero_grad()
    model_inputs = batch["model_inputs"]
    device = next(diffuser.denoiser.parameters()).device
    timesteps = diffuser.draw_timesteps(model_inputs["x"].shape[0]).to(device)
    model_inputs.update({"p": p_classifier_free_guidance})
    losses = diffuser.compute_loss(model_inputs=model_inputs, timesteps=timesteps, extra_args=batch.get("extra", {}))
    loss = sum(losses.values())
    if reducer is not None:
        reducer.begin()
    loss.backward()
    if reducer is not None:
        reducer.finish()
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return {k: v.detach() for k, v in losses.items()}
