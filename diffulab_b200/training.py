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

_SYNC_ZERO = bool(__import__("os").environ.get("DLB_SYNC_ZERO"))  # A/B switch: gradient memset on the compute stream
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

    def use_grad_buffer(self, buf: Tensor) -> None:
        """Move the flat gradient buffer (e.g. into symmetric / peer-mapped memory for the NVLink gradient reduction)."""
        assert buf.dtype == torch.float32 and buf.numel() == self.total and buf.device == self.flat_g.device and buf.is_contiguous()
        buf.copy_(self.flat_g)
        self.flat_g = buf
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat_g[o : o + p.numel()].view(p.shape)

    def zero_grad(self) -> None:
        self.flat_g.zero_()
        for p, o in zip(self.params, self.offsets):  # re-attach if someone set .grad to None
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o : o + p.numel()].view(p.shape)


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction in double; no amsgrad) as ONE kernel over the
    flat parameter buffer; the same pass writes the bf16 shadow consumed by the next forward and, when an `EMA` is
    attached and due, its moving average (reference: torch.optim.AdamW built from configs/optimizer/adamw.yaml +
    ema_pytorch, base_trainer.py:149-153). Parameters that received no gradient since `zero_grad()` are skipped exactly as
    torch skips `p.grad is None` (no decay, no moment update). `state_dict()` / `load_state_dict()` keep torch's
    per-parameter layout (step / exp_avg / exp_avg_sq)."""

    def __init__(self, params, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, grad_scale: float = 1.0):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_scale = grad_scale
        self.stores: list[FlatParams] = []
        self._steps: list[int] = []
        self._m: list[Tensor] = []
        self._v: list[Tensor] = []
        self._step_t: list[Tensor] = []
        self._mask_key: list[frozenset | None] = []
        self._mask: list[Tensor | None] = []
        self.ema: "EMA | None" = None
        self._zero_stream: Any = None
        self._zero_done: Any = None
        for group in self.param_groups:
            st = FlatParams(group["params"])
            self.stores.append(st)
            self._steps.append(0)
            self._m.append(torch.zeros_like(st.flat_p))
            self._v.append(torch.zeros_like(st.flat_p))
            self._step_t.append(torch.tensor(0.0))  # one shared counter tensor per group (torch keeps one per parameter)
            self._mask_key.append(None)
            self._mask.append(None)
        self._point_state()

    def _point_state(self) -> None:
        """self.state[p] = views of the flat moment buffers (what state_dict() serialises)."""
        for gi, st in enumerate(self.stores):
            m, v = self._m[gi], self._v[gi]
            for p, o in zip(st.params, st.offsets):
                n = p.numel()
                self.state[p] = {"step": self._step_t[gi], "exp_avg": m[o : o + n].view(p.shape), "exp_avg_sq": v[o : o + n].view(p.shape)}

    def load_state_dict(self, state_dict: dict) -> None:
        """torch's loader replaces self.state[p] with fresh tensors; copy them back into the flat buffers the kernel reads,
        restore the step counters and re-point the state at the flat views (a resumed run continues bit-identically)."""
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for gi, st in enumerate(self.stores):
                steps = set()
                for p, o in zip(st.params, st.offsets):
                    ent = self.state.get(p, {})
                    n = p.numel()
                    if "exp_avg" in ent:
                        self._m[gi][o : o + n].copy_(ent["exp_avg"].reshape(-1))
                        self._v[gi][o : o + n].copy_(ent["exp_avg_sq"].reshape(-1))
                    if "step" in ent:
                        steps.add(int(float(ent["step"])))
                step = max(steps) if steps else 0
                self._steps[gi] = step
                self._step_t[gi] = torch.tensor(float(step))
        self._point_state()

    def zero_grad(self, set_to_none: bool = True) -> None:  # noqa: ARG002 - gradients live in the flat buffer
        self.wait_zero()
        for st in self.stores:
            st.zero_grad()
        K.clear_touched()

    def zero_grad_async(self) -> None:
        """zero_grad() on a side stream: the 3.3 GB memset of the flat gradient buffer (0.5 ms of HBM time for DiT-XL/2) overlaps
        the forward pass, which never touches gradients. The caller must call `wait_zero()` before the first gradient is written
        (`training_step` does, right before `backward()`)."""
        cur = torch.cuda.current_stream()
        if self._zero_stream is None:
            self._zero_stream = torch.cuda.Stream()
        start = torch.cuda.Event()
        start.record(cur)  # everything that read the gradients (optimizer step, peers' pulls) is ordered before this point
        with torch.cuda.stream(self._zero_stream):
            self._zero_stream.wait_event(start)
            for st in self.stores:
                st.zero_grad()
            self._zero_done = torch.cuda.Event()
            self._zero_done.record()
        K.clear_touched()

    def wait_zero(self) -> None:
        if self._zero_done is not None:
            torch.cuda.current_stream().wait_event(self._zero_done)
            self._zero_done = None

    def _active_mask(self, gi: int) -> Tensor | None:
        """uint8 per 64-element chunk: 0 for parameters that received no gradient this step (torch: .grad is None).
        Rebuilt only when the set of gradient-receiving parameters changes (normally once)."""
        st = self.stores[gi]
        touched = frozenset(i for i, p in enumerate(st.params) if id(p) in K._touched)
        if touched == self._mask_key[gi]:
            return self._mask[gi]
        self._mask_key[gi] = touched
        idle = [i for i in range(len(st.params)) if i not in touched]
        # a gradient may also have arrived through plain autograd (accumulated into the flat view): only all-zero slices
        # of untouched parameters count as "no gradient"
        idle = [i for i in idle if not bool(st.params[i].grad.any().item())] if idle else []
        if not idle:
            self._mask[gi] = None
            return None
        mask = torch.ones(st.total // _ALIGN, dtype=torch.uint8)
        for i in idle:
            o, n = st.offsets[i], st.params[i].numel()
            mask[o // _ALIGN : (o + n + _ALIGN - 1) // _ALIGN] = 0
        self._mask[gi] = mask.to(st.flat_p.device)
        return self._mask[gi]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.wait_zero()
        ema_plan = self.ema.plan() if self.ema is not None else None  # ("skip" | "copy" | "lerp", decay)
        for gi, (group, st) in enumerate(zip(self.param_groups, self.stores)):
            st.check_versions()
            self._steps[gi] += 1
            b1, b2 = group["betas"]
            mask = self._active_mask(gi)
            # the EMA rides in the same pass only when every parameter is written (no skipped chunks)
            fuse_ema = ema_plan is not None and ema_plan[0] != "skip" and mask is None
            ops.adamw_step(st.flat_p, st.flat_g, self._m[gi], self._v[gi], st.flat_shadow, lr=float(group["lr"]), beta1=b1, beta2=b2,
                           eps=group["eps"], weight_decay=group["weight_decay"], step=self._steps[gi], grad_scale=self.grad_scale,
                           ema=self.ema.flat[gi] if fuse_ema else None, ema_decay=ema_plan[1] if fuse_ema else 0.0, chunk_active=mask)
            if ema_plan is not None and ema_plan[0] != "skip" and not fuse_ema:
                ops.ema_lerp_(self.ema.flat[gi], st.flat_p, ema_plan[1])
            self._step_t[gi] += 1
        if self.ema is not None:
            self.ema.advance()
        K.bump_shadow_epoch()  # version-cached (padded) shadows must be rebuilt; installed flat shadows stay valid
        return loss


class EMA:
    """Exponential moving average of the optimised parameters, fused into the AdamW pass.

    Reference: `ema_pytorch.EMA(model, beta, update_after_step, update_every)` built in base_trainer.py:247-253 and
    advanced by `ema_denoiser.update()` right after `optimizer.step()` (base_trainer.py:152-153). ema-pytorch 0.7.7 is a
    third-party dependency (uv.lock) that is NOT present in this container, so its published algorithm is restated here
    ("parity unpinned" for this class; tests compare with a torch restatement of the same rule):
      update(): step = self.step; self.step += 1
                first call            -> copy the online parameters (initted)
                step % update_every   != 0 -> nothing
                step <= update_after_step  -> copy the online parameters
                else                  -> ema.lerp_(online, 1 - decay(step+1))
      decay(s): e = max(s - update_after_step - 1, 0); 0 if e <= 0 else clamp(1 - (1 + e / inv_gamma) ** -power, min_value, beta)
    with the package defaults inv_gamma = 1.0, power = 2/3, min_value = 0.0.
    Only trainable parameters are averaged (the denoisers on this path have no buffers)."""

    def __init__(self, optimizer: FusedAdamW, beta: float = 0.9999, update_after_step: int = 100, update_every: int = 10,
                 inv_gamma: float = 1.0, power: float = 2.0 / 3.0, min_value: float = 0.0):
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.step = 0
        self.initted = False
        self.opt = optimizer
        self.flat = [st.flat_p.clone() for st in optimizer.stores]
        optimizer.ema = self

    def get_current_decay(self, step: int | None = None) -> float:
        s = self.step if step is None else step
        epoch = max(s - self.update_after_step - 1, 0)
        if epoch <= 0:
            return 0.0
        value = 1.0 - (1.0 + epoch / self.inv_gamma) ** (-self.power)
        return min(max(value, self.min_value), self.beta)

    def plan(self) -> tuple[str, float]:
        """What the update following the current optimizer step does: ('skip' | 'copy' | 'lerp', decay)."""
        step = self.step
        if not self.initted:
            return ("copy", 0.0)
        if step % self.update_every != 0:
            return ("skip", 0.0)
        if step <= self.update_after_step:
            return ("copy", 0.0)
        return ("lerp", self.get_current_decay(step + 1))

    def advance(self) -> None:
        self.step += 1
        self.initted = True

    def update(self) -> None:
        """Stand-alone update (when the optimizer step was not the fused one): same rule, separate launch."""
        kind, decay = self.plan()
        if kind != "skip":
            for e, st in zip(self.flat, self.opt.stores):
                ops.ema_lerp_(e, st.flat_p, decay)
        self.advance()

    def state_dict(self) -> dict[str, Tensor]:
        """EMA weights under the model's parameter names order (list index -> tensor), fp32."""
        out = {}
        for gi, st in enumerate(self.opt.stores):
            for i, (p, o) in enumerate(zip(st.params, st.offsets)):
                out[f"{gi}.{i}"] = self.flat[gi][o : o + p.numel()].view(p.shape)
        return out

    def parameters_like(self, model: nn.Module) -> dict[str, Tensor]:
        """EMA tensors keyed by `model`'s parameter names (what `ema_model.state_dict()` holds in the reference)."""
        by_id = {}
        for gi, st in enumerate(self.opt.stores):
            for p, o in zip(st.params, st.offsets):
                by_id[id(p)] = self.flat[gi][o : o + p.numel()].view(p.shape)
        return {n: by_id[id(p)] for n, p in model.named_parameters() if id(p) in by_id}


def make_reduce_group(comm_ctas: int = 4) -> Any:
    """A dedicated NCCL communicator for the gradient buckets, limited to `comm_ctas` CTAs per collective. The buckets need
    ~3.3 GB per ~95 ms of backward (DiT-XL/2, fp32) = 35 GB/s — a few CTAs over NVLink 5 — while NCCL's default of 16-32
    CTAs would time-slice with as many persistent GEMM CTAs and make each of them the straggler of its GEMM. Returns None
    (default group) when the job is not distributed or the backend is not NCCL."""
    if not dist.is_initialized() or dist.get_backend() != "nccl" or comm_ctas <= 0:
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = int(comm_ctas)
    opts.config.min_ctas = min(int(comm_ctas), 2)
    return dist.new_group(backend="nccl", pg_options=opts)


class GradReducer:
    """Bucketed gradient all-reduce (mean over ranks) launched during backward.

    Buckets are contiguous ranges of a FlatParams gradient buffer (or per-parameter tensors when no flat store is
    given). `blocks.set_grad_ready_hook` reports each parameter the moment its gradient kernels have been enqueued;
    when all parameters of a bucket are ready the bucket's all-reduce is issued with async_op=True (c10d orders it
    after the compute stream's current position and runs it on the communicator's own stream), so communication
    overlaps the rest of backward. Parameters that never receive a gradient (reference quirk) are
    tolerated: their buckets are flushed, in a fixed order, by `finish()`.

    `tail_bucket_mb`: the buckets produced LAST by backward (the first-registered parameters) are capped at this size so that
    the part of the reduction that cannot overlap anything is short. `reserve_sms`: between `begin()` and `finish()` the
    persistent GEMM / attention kernels leave this many SMs to the collective (dlb_set_sm_budget).

    `mode`:
      "nccl"  ncclAllReduce(AVG) per bucket on the communicator's stream (16-32 CTAs that time-slice with the persistent GEMMs:
              measured +5 ms of kernel time per 143 ms step at N = 2, profiles/dp_r2/README.md).
      "ce"    the flat gradient buffer lives in symmetric (peer-mapped) memory; per bucket, rank p pulls piece p of every peer with
              device-to-device copies (copy engines over NVLink, no SMs), averages them (dlb_reduce_pieces), and all ranks pull
              the reduced pieces back. Stream-ordered symmetric-memory barriers order the phases across ranks.
      "nvls"  one dlb_multimem_allreduce kernel per bucket (sum inside the NVSwitch, `comm_ctas` CTAs) after a barrier; the closing
              barrier of `finish()` orders the peers' multicast stores before this rank reads its buffer.
    Modes "ce" / "nvls" need a FlatParams store and an NCCL (CUDA) job; the mean is exact in fp32 as with NCCL."""

    DEFAULT_BUCKET_MB = 128.0

    def __init__(self, stores: list[FlatParams] | None = None, params: list[nn.Parameter] | None = None,
                 bucket_mb: float = DEFAULT_BUCKET_MB, process_group: Any = None, tail_bucket_mb: float | None = None,
                 reserve_sms: int = 0, mode: str = "nccl", comm_ctas: int = 4, use_graphs: bool = True):
        assert mode in ("nccl", "ce", "nvls"), mode
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.reserve_sms = int(reserve_sms)
        self.mode = mode if self.world > 1 else "nccl"
        self.comm_ctas = int(comm_ctas)
        self.bucket_span: list[tuple[int, int, int]] = []  # (store index, start, end) of every bucket in its flat buffer
        self._sym: list[Any] = []
        if self.mode != "nccl":
            assert stores and all(st.flat_g.is_cuda for st in stores), "peer-memory gradient reduction needs FlatParams stores on CUDA"
            import torch.distributed._symmetric_memory as symm_mem

            group = process_group if process_group is not None else dist.group.WORLD
            self.rank = dist.get_rank(group)
            for st in stores:
                buf = symm_mem.empty(st.total, dtype=torch.float32, device=st.flat_g.device)
                st.use_grad_buffer(buf)
                self._sym.append(symm_mem.rendezvous(buf, group))
            if self.mode == "nvls":
                assert all(int(h.multicast_ptr) != 0 for h in self._sym), "no NVLink multicast (NVLS) mapping on this system: use mode='ce'"
            self.comm = torch.cuda.Stream(priority=-1)
            self._stage: Tensor | None = None
            self._plans: dict[int, dict[str, Any]] = {}
            self._graphs: dict[int, Any] = {}
            self._uses: dict[int, int] = {}
            self.use_graphs = use_graphs
        self.buckets: list[Tensor] = []
        self.bucket_of: dict[int, int] = {}
        self.pending_init: list[int] = []
        self.timeline: list[tuple[int, Any]] | None = None
        cap = int(bucket_mb * 1024 * 1024 / 4)
        tail_cap = int((tail_bucket_mb if tail_bucket_mb else bucket_mb) * 1024 * 1024 / 4)
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
                    # what is still ahead of this parameter (offsets below `o`) is produced later in backward: once less than
                    # two tail buckets remain, switch to the small cap
                    if count >= (tail_cap if o <= 2 * tail_cap else cap):
                        self._add_bucket(st.flat_g[start:end], members)
                        self.bucket_span.append((stores.index(st), start, end))
                        start, count, members = None, 0, []
                if members:
                    self._add_bucket(st.flat_g[start:end], members)
                    self.bucket_span.append((stores.index(st), start, end))
        else:
            assert params is not None
            for p in reversed([q for q in params if q.requires_grad]):
                self._add_bucket(K.gbuf(p).view(-1), [p])
        self.pending = list(self.pending_init)
        self.launched = [False] * len(self.buckets)
        self.works: list[Any] = []
        self.active = False
        if self.mode == "ce":  # staging for the world - 1 peer copies of the largest piece (shared by all buckets: one stream)
            piece_max = max(self._pieces(e - s_)[0] for _, s_, e in self.bucket_span)
            self._stage = torch.empty((self.world - 1) * piece_max, device=self.buckets[0].device, dtype=torch.float32)

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
        self.seen: set[int] = set()
        self.active = True
        K.set_grad_ready_hook(self._ready)
        if self.reserve_sms > 0 and self.world > 1:
            from . import _lib
            _lib.set_sm_budget(torch.cuda.get_device_properties(self.buckets[0].device).multi_processor_count - self.reserve_sms)

    def _launch(self, bi: int) -> None:
        if self.launched[bi]:
            return
        self.launched[bi] = True
        if self.world == 1:
            return
        t = self.buckets[bi]
        if t.is_cuda:
            if self.timeline is not None:  # where on the compute stream this bucket became ready
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                self.timeline.append((bi, ev))
            if self.mode == "nccl":
                self.works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
            else:
                self.works.append(self._peer_reduce(bi))
        else:  # gloo (CPU tests of the host logic) has no AVG
            w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self.works.append((w, t))

    _BARRIER_TIMEOUT_MS = 30000  # a lost peer traps the kernel instead of hanging the GPU

    def _pieces(self, n: int) -> tuple[int, list[int]]:
        """piece length (multiple of 64 floats) and per-rank lengths of a bucket of n floats (the last pieces may be short / empty)"""
        piece = ((n + self.world - 1) // self.world + _ALIGN - 1) // _ALIGN * _ALIGN
        return piece, [max(0, min(piece, n - q * piece)) for q in range(self.world)]

    def _plan(self, bi: int) -> dict[str, Any]:
        """Views for bucket bi, built once: my piece, (staging slot <- peer's copy of my piece) pulls, (my copy of a peer's piece <-
        the peer's reduced piece) pulls. Peer views come from the symmetric-memory handle and never change."""
        if bi in self._plans:
            return self._plans[bi]
        si, start, end = self.bucket_span[bi]
        hdl, t = self._sym[si], self.buckets[bi]
        W, r = self.world, self.rank
        piece, lens = self._pieces(end - start)
        plan: dict[str, Any] = {"hdl": hdl, "piece": piece, "len": lens[r], "own": t[r * piece : r * piece + lens[r]], "pull": [], "gather": [],
                                "mc": (int(hdl.multicast_ptr) + 4 * (start + r * piece)) if self.mode == "nvls" else 0}
        for s_ in range(W - 1):
            peer = (r + 1 + s_) % W
            if lens[r] > 0 and self.mode == "ce":
                plan["pull"].append((self._stage[s_ * piece : s_ * piece + lens[r]], hdl.get_buffer(peer, (lens[r],), torch.float32, start + r * piece)))
            if lens[peer] > 0 and self.mode == "ce":
                plan["gather"].append((t[peer * piece : peer * piece + lens[peer]], hdl.get_buffer(peer, (lens[peer],), torch.float32, start + peer * piece)))
        self._plans[bi] = plan
        return plan

    def _enqueue(self, bi: int) -> None:
        """The reduction of bucket bi on the CURRENT stream (the communication stream, possibly under graph capture)."""
        pl = self._plan(bi)
        hdl, W = pl["hdl"], self.world
        hdl.barrier(channel=0, timeout_ms=self._BARRIER_TIMEOUT_MS)  # every rank's gradients of this bucket are complete
        if self.mode == "nvls":
            if pl["len"] > 0:
                ops.multimem_allreduce_(pl["mc"], pl["len"], 1.0 / W, self.comm_ctas)
            return
        for dst, src in pl["pull"]:  # my piece of every peer's bucket (copy engines over NVLink)
            dst.copy_(src, non_blocking=True)
        if pl["len"] > 0:
            ops.reduce_pieces_(pl["own"], self._stage, pl["piece"], W, self.rank, 1.0 / W)
        hdl.barrier(channel=0, timeout_ms=self._BARRIER_TIMEOUT_MS)  # every rank's piece is reduced
        for dst, src in pl["gather"]:  # the reduced pieces of the peers (copy engines again)
            dst.copy_(src, non_blocking=True)

    def _peer_reduce(self, bi: int) -> Any:
        """Reduce bucket bi over NVLink peer memory on the communication stream; returns an event the compute stream waits for.
        The per-bucket sequence (2 barriers, 2 (W - 1) copies, 1 kernel: 17 launches at W = 8) is recorded ONCE as a CUDA graph
        (second use; the first runs eagerly) and replayed with one launch afterwards — at W = 8 the eager launches cost the
        backward thread ~5 ms per step, which starved the compute stream."""
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            graph = self._graphs.get(bi)
            if graph is not None:
                graph.replay()
            elif self.use_graphs and self._uses.get(bi, 0) >= 1:
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.comm, capture_error_mode="thread_local"):
                        self._enqueue(bi)
                    self._graphs[bi] = g
                    g.replay()
                except Exception as e:  # noqa: BLE001 - capture is an optimisation: fall back to eager launches, loudly
                    import warnings

                    warnings.warn(f"GradReducer: CUDA-graph capture of the bucket reduction failed ({e!r}); launching eagerly", stacklevel=2)
                    self.use_graphs = False
                    self._enqueue(bi)
            else:
                self._enqueue(bi)
            self._uses[bi] = self._uses.get(bi, 0) + 1
            done = torch.cuda.Event()
            done.record()
        return done

    def _ready(self, p: Tensor) -> None:
        if not self.active:
            return
        bi = self.bucket_of.get(id(p))
        if bi is None:
            return
        # `_ready` fires once per USE of a parameter; a parameter used twice in one forward (shared module, tied weight)
        # would release its bucket before the second wgrad was enqueued. Count each parameter once and refuse reuse after
        # its bucket is in flight (ranks would silently diverge otherwise).
        if id(p) in self.seen:
            if self.launched[bi]:
                raise RuntimeError("GradReducer: a parameter received a second gradient after its bucket was all-reduced "
                                   "(shared / tied weights are not supported by the overlapped reducer)")
            return
        self.seen.add(id(p))
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def finish(self) -> None:
        """Flush buckets holding never-produced gradients (fixed order on every rank), then wait for all of them."""
        K.set_grad_ready_hook(None)
        self.active = False
        if self.reserve_sms > 0 and self.world > 1:
            from . import _lib
            _lib.set_sm_budget(0)
        for bi in range(len(self.buckets)):
            self._launch(bi)
        done: list[Any] = []
        if self.timeline is not None and self.works and not isinstance(self.works[0], tuple):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timeline.append((-1, ev))  # end of backward on the compute stream
        if self.mode != "nccl" and self.works:
            # peers read my pieces (and, multicast mode, write theirs into my buffer) until THEIR last operation: one closing barrier
            # per store before this rank's optimizer reads / the next zero_grad overwrites the gradient buffer
            with torch.cuda.stream(self.comm):
                for hdl in self._sym:
                    hdl.barrier(channel=0, timeout_ms=self._BARRIER_TIMEOUT_MS)
                closing = torch.cuda.Event()
                closing.record()
        for w in self.works:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world)
            else:
                if self.mode == "nccl":
                    w.wait()
                else:
                    torch.cuda.current_stream().wait_event(w)
                if self.timeline is not None:  # the compute stream now waits for this bucket: an event here is its completion
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record()
                    done.append(ev)
        if self.mode != "nccl" and self.works:
            torch.cuda.current_stream().wait_event(closing)
        if self.timeline is not None:
            self.timeline.extend((-2 - i, ev) for i, ev in enumerate(done))
        self.works = []

    def read_timeline(self) -> dict[str, Any]:
        """After a step run with `self.timeline = []` (and a device synchronize): per-bucket ready / done times in ms relative
        to the first bucket's ready point, the end of backward, and the exposed tail = last completion - end of backward."""
        assert self.timeline, "set reducer.timeline = [] before the step"
        t0 = self.timeline[0][1]
        ready = {bi: t0.elapsed_time(ev) for bi, ev in self.timeline if bi >= 0}
        end_bwd = next(t0.elapsed_time(ev) for bi, ev in self.timeline if bi == -1)
        done = [t0.elapsed_time(ev) for bi, ev in self.timeline if bi <= -2]
        order = [bi for bi, _ in self.timeline if bi >= 0]
        return {"bucket_mb": [round(self.buckets[bi].numel() * 4 / 2**20, 1) for bi in order], "ready_ms": [round(ready[bi], 3) for bi in order],
                "done_ms": [round(d, 3) for d in done], "end_backward_ms": round(end_bwd, 3),
                "exposed_tail_ms": round(max(done) - end_bwd, 3) if done else 0.0}


class DevicePrefetcher:
    """Iterates host batches (pinned memory) as device batches, copying batch i + 1 on a side stream while step i computes —
    the place of `accelerator.prepare(train_dataloader)` in the reference's loop (base_trainer.py:277-307: the prepared DataLoader
    moves every batch to the device). `restructure(host_batch)` (optional) reshapes the host batch into the structure the step
    takes WITHOUT copying (nested dicts of tensors; anything else passes through). The leaves are copied into TWO persistent sets
    of device buffers used alternately (no allocation in the loop: freeing cross-stream blocks every step makes the caching
    allocator fall back to cudaMalloc, which serialises host and device). Ordering: the compute stream waits for the copy's event;
    the copy into a buffer set waits until the step that last used it has been fully enqueued and executed. A batch must not be
    used after the NEXT batch has been requested twice (it is overwritten)."""

    _streams: dict = {}
    N_BUF = 2

    def __init__(self, batches: Iterable[Any], device: torch.device, restructure: Any = None):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.restructure = restructure or (lambda b: b)
        key = (self.device.type, self.device.index if self.device.index is not None else torch.cuda.current_device())
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=self.device)
        self.stream = DevicePrefetcher._streams[key]
        self._bufs: list[Any] = [None] * self.N_BUF        # persistent device leaves, same nesting as the host batch
        self._reusable: list[Any] = [None] * self.N_BUF    # compute-stream event after which the set may be overwritten
        self._n = 0                                        # batches issued so far
        self._next: tuple[Any, torch.cuda.Event] | None = None
        self._fill()

    def _copy(self, dst: Any, src: Any) -> tuple[Any, Any]:
        """-> (persistent structure, view handed to the consumer); tensors are copied on the current (side) stream"""
        if isinstance(src, Tensor):
            if not isinstance(dst, Tensor) or dst.shape != src.shape or dst.dtype != src.dtype:
                dst = torch.empty(src.shape, dtype=src.dtype, device=self.device)
            dst.copy_(src, non_blocking=True)
            return dst, dst
        if isinstance(src, dict):
            keep, out = {}, {}
            for k, v in src.items():
                keep[k], out[k] = self._copy(dst.get(k) if isinstance(dst, dict) else None, v)
            return keep, out
        return src, src

    def _fill(self) -> None:
        try:
            host = self.restructure(next(self.it))
        except StopIteration:
            self._next = None
            return
        k = self._n % self.N_BUF
        self._n += 1
        with torch.cuda.stream(self.stream):
            if self._reusable[k] is not None:
                self.stream.wait_event(self._reusable[k])
            self._bufs[k], dev = self._copy(self._bufs[k], host)
            ev = torch.cuda.Event()
            ev.record()
        self._next = (dev, ev)

    def __iter__(self):
        return self

    def __next__(self) -> Any:
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        # everything enqueued so far belongs to steps that used EARLIER batches: the set holding the previous batch may be refilled
        # once the compute stream has passed this point
        if self._n >= 2:
            done = torch.cuda.Event()
            done.record(cur)
            self._reusable[(self._n - 2) % self.N_BUF] = done
        cur.wait_event(ev)
        self._fill()  # the next batch's copy overlaps the step the caller is about to run
        return dev


class LossReader:
    """Per-step device -> host read of the loss dict without stalling the launch queue: `push(losses)` starts an asynchronous
    copy into pinned memory and returns the PREVIOUS step's values (None on the first call); `flush()` returns the last one.
    (The reference calls `loss.item()` inside the step, base_trainer.py:125-137, which serialises host and device every step.)"""

    def __init__(self):
        self._pending: tuple[list[str], Tensor, torch.cuda.Event] | None = None
        self._host: list[Tensor] = []
        self._i = 0

    def _take(self) -> dict[str, float] | None:
        if self._pending is None:
            return None
        keys, host, ev = self._pending
        ev.synchronize()
        self._pending = None
        return {k: float(host[i]) for i, k in enumerate(keys)}

    def push(self, losses: dict[str, Tensor]) -> dict[str, float] | None:
        keys = list(losses)
        dev = torch.stack([losses[k].detach().float().reshape(()) for k in keys])
        if len(self._host) < 2 or self._host[0].numel() != len(keys):
            self._host = [torch.empty(len(keys), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        prev = self._take()
        host = self._host[self._i]
        self._i ^= 1
        host.copy_(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending = (keys, host, ev)
        return prev

    def flush(self) -> dict[str, float] | None:
        return self._take()


def training_step(diffuser, optimizer: torch.optim.Optimizer, batch: dict[str, Any], p_classifier_free_guidance: float = 0.0,
                  reducer: GradReducer | None = None, scheduler: Any | None = None, per_batch_scheduler: bool = False,
                  ema: EMA | None = None) -> dict[str, Tensor]:
    """One optimisation step with the reference's order of operations (base_trainer.py:138-153): zero_grad -> draw t ->
    set p -> compute_loss -> backward (+ bucketed all-reduce) -> optimizer.step -> scheduler.step if per_batch_scheduler ->
    EMA update. An `EMA` attached to a `FusedAdamW` is advanced inside `optimizer.step()` (same kernel pass); any other
    combination calls `ema.update()` here. Returns the loss dict as DEVICE tensors: the reference's per-step
    `loss.item()` host sync is left to the caller."""
    if hasattr(optimizer, "zero_grad_async") and not _SYNC_ZERO:
        optimizer.zero_grad_async()  # the gradient memset runs beside the forward pass
    else:
        optimizer.zero_grad()
    model_inputs = batch["model_inputs"]
    device = next(diffuser.denoiser.parameters()).device
    timesteps = diffuser.draw_timesteps(model_inputs["x"].shape[0]).to(device)
    model_inputs.update({"p": p_classifier_free_guidance})
    losses = diffuser.compute_loss(model_inputs=model_inputs, timesteps=timesteps, extra_args=batch.get("extra", {}))
    loss = sum(losses.values())
    if hasattr(optimizer, "wait_zero"):
        optimizer.wait_zero()
    if reducer is not None:
        reducer.begin()
    loss.backward()
    if reducer is not None:
        reducer.finish()
    optimizer.step()
    if scheduler is not None and per_batch_scheduler:
        scheduler.step()
    if ema is not None and getattr(optimizer, "ema", None) is not ema:
        ema.update()
    return {k: v.detach() for k, v in losses.items()}
