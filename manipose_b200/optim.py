"""Flat fp32 parameter / gradient storage, fused Adam and the bucketed data-parallel gradient reducer.

Reference: ``torch.optim.Adam(model.parameters(), lr=4e-5, weight_decay=1e-6)`` (hpe/main_h36m_lifting.py:755-761) under
``nn.DataParallel``.  Here (training): one process per GPU; every parameter and its gradient are views into
two flat fp32 buffers, so that

* the backward sweep accumulates weight gradients in place (``train_ops.grad_of``),
* gradient averaging is a handful of NCCL all-reduces over contiguous buckets (one per transformer block, launched from
  the backward sweep as soon as a block's gradients are complete, so they overlap the rest of the backward), and
* the optimizer step is ONE kernel (``mp_adam_step``) over the flat buffers.

``FusedAdam`` follows ``torch.optim.Adam`` numerics (L2-style weight decay, bias correction) and exposes ``param_groups`` so
``ReduceLROnPlateau`` (main_h36m_lifting.py:763-771) works unchanged.  ``FlatParameters`` and ``GradientReducer`` are device
agnostic (the world-size-2 gloo tests run them on CPU); only ``FusedAdam.step`` needs the sm_100a library.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

_ALIGN = 64   # floats: every parameter starts on a 256-byte boundary (TMA / vector alignment of the gradient GEMMs)


class FlatParameters:
    """Re-homes the parameters of ``module`` (already on their final device) into one flat buffer; ``p.grad`` become views of a
    second one.  ``state_dict`` / ``load_state_dict`` keep working (they copy in place)."""

    def __init__(self, module: nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev = params[0].device
        self.params = params
        self.offsets: Dict[int, Tuple[int, int]] = {}
        off = 0
        for p in params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("FlatParameters needs fp32 parameters on one device")
            self.offsets[id(p)] = (off, p.numel())
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.flat_param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in params:
                o, n = self.offsets[id(p)]
                view = self.flat_param[o:o + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.attach_grads()

    def attach_grads(self) -> None:
        for p in self.params:
            o, n = self.offsets[id(p)]
            g = self.flat_grad[o:o + n].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g

    def zero_grad(self) -> None:
        self.flat_grad.zero_()
        self.attach_grads()

    def span(self, params: Sequence[torch.Tensor]) -> Tuple[int, int]:
        """[start, stop) of the flat buffers covering ``params`` (they must be contiguous in registration order)."""
        spans = sorted(self.offsets[id(p)] for p in params)
        start = spans[0][0]
        stop = spans[-1][0] + (spans[-1][1] + _ALIGN - 1) // _ALIGN * _ALIGN
        return start, min(stop, self.numel)


class GradientReducer:
    """Sum-all-reduce of ``flat_grad`` in contiguous buckets.  ``bucket_ready(i)`` launches bucket i asynchronously (NCCL runs it on
    its own stream after the work already queued on the current stream); ``finish()`` launches whatever was not reduced yet and
    waits.  The 1 / world_size averaging is folded into the optimizer step (``FusedAdam.step(grad_scale=...)``)."""

    def __init__(self, flat_grad: torch.Tensor, buckets: List[Tuple[int, int]], group=None):
        self.flat_grad = flat_grad
        self.buckets = list(buckets)
        self.group = group
        self._handles = []
        self._done = [False] * len(self.buckets)
        covered = sorted(self.buckets)
        if covered[0][0] != 0 or covered[-1][1] != flat_grad.numel() or any(a[1] != b[0] for a, b in zip(covered, covered[1:])):
            raise ValueError("buckets must tile the flat gradient buffer")

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def bucket_ready(self, i: int) -> None:
        if self._done[i]:
            raise RuntimeError(f"gradient bucket {i} reduced twice in one step: two backward passes before one optimizer.step() "
                               "(gradient accumulation) are not supported with overlap_reduce=True; build FusedAdam(..., "
                               "overlap_reduce=False) to reduce once, in step()")
        self._done[i] = True
        if self.world_size > 1:
            a, b = self.buckets[i]
            self._handles.append(dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> float:
        """Reduces the remaining buckets, waits for all of them, returns the factor that turns the sums into means."""
        for i, done in enumerate(self._done):
            if not done:
                self.bucket_ready(i)
        for h in self._handles:
            h.wait()
        self._handles = []
        self._done = [False] * len(self.buckets)
        return 1.0 / self.world_size


def block_buckets(flat: FlatParameters, model: nn.Module) -> Tuple[List[Tuple[int, int]], Dict[int, int]]:
    """One bucket per transformer block of the rotations backbone (2.1 M parameters each) plus the gaps between them (embeddings,
    shared norms, heads, the small bone-length backbone).  Returns (buckets, {id(block): bucket index})."""
    spans = []
    rot = getattr(model, "rotations_module", model)
    for blk in list(getattr(rot, "STEblocks", [])) + list(getattr(rot, "TTEblocks", [])):
        ps = [p for p in blk.parameters() if p.requires_grad]
        if ps:
            spans.append((flat.span(ps), id(blk)))
    spans.sort()
    buckets, index, cursor = [], {}, 0
    for (a, b), key in spans:
        if a > cursor:
            buckets.append((cursor, a))
        index[key] = len(buckets)
        buckets.append((a, b))
        cursor = b
    if cursor < flat.numel:
        buckets.append((cursor, flat.numel))
    return buckets, index


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` (no amsgrad) as one kernel over flat parameter / gradient / moment buffers."""

    def __init__(self, module: nn.Module, lr: float = 4e-5, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, group=None, overlap_reduce: bool = True):
        self._module = module
        self.flat = FlatParameters(module)
        super().__init__(self.flat.params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.exp_avg = torch.zeros_like(self.flat.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat.flat_param)
        self.step_count = 0
        # the authoritative step count lives on the device so that a CUDA-graph capture of the training step advances it on replay
        self.step_dev = torch.zeros((), dtype=torch.int64, device=self.flat.flat_param.device)
        # ... and so does the learning rate: a scheduler writes param_groups[0]["lr"] on the host, ``_sync_lr`` copies it over
        # before the kernel (eager) / before each replay (``CapturedTrainStep``), the kernel reads the device value
        self.lr_dev = torch.full((), float(lr), dtype=torch.float32, device=self.flat.flat_param.device)
        self._lr_on_dev = float(lr)
        buckets, index = block_buckets(self.flat, module)
        self.reducer = GradientReducer(self.flat.flat_grad, buckets, group)
        self._bucket_of_block = index
        if overlap_reduce:
            rot = getattr(module, "rotations_module", None)
            if rot is not None:
                rot._on_block_grads = self._block_done
        self.broadcast_state()

    def broadcast_state(self, src: int = 0) -> None:
        """Data parallel: every replica starts from rank ``src``'s parameters, moments and step count (what torch DDP does at
        construction; the reference's nn.DataParallel has a single copy).  Called from the constructor and ``load_state_dict``, so
        the replicas agree even when the ranks were seeded differently or only one of them read the checkpoint."""
        if self.reducer.world_size <= 1:
            return
        group = self.reducer.group
        src_global = dist.get_global_rank(group, src) if group is not None else src
        for t in (self.flat.flat_param, self.exp_avg, self.exp_avg_sq):
            dist.broadcast(t, src=src_global, group=group)
        step = self.step_dev.reshape(1).clone()
        dist.broadcast(step, src=src_global, group=group)
        self.step_dev.copy_(step[0])
        self.step_count = int(step[0].item())
        self._invalidate_shadows()

    def _sync_lr(self) -> None:
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_dev:
            self.lr_dev.fill_(lr)
            self._lr_on_dev = lr

    def _check_homes(self) -> None:
        """The parameters must still be views of the flat buffer: model.to() / .half() / load_state_dict(assign=True) after the
        optimizer was built re-home them, and the kernel would then update a buffer the model no longer reads."""
        base, nbytes = self.flat.flat_param.data_ptr(), self.flat.flat_param.numel() * 4
        for p in (self.flat.params[0], self.flat.params[-1]):
            if not (base <= p.data_ptr() < base + nbytes):
                raise RuntimeError("FusedAdam: the model's parameters no longer live in the optimizer's flat buffer (model.to() / .half() / "
                                   "load_state_dict(assign=True) after FusedAdam(model)?); build the optimizer after moving the model")

    def _block_done(self, blk: nn.Module) -> None:
        i = self._bucket_of_block.get(id(blk))
        if i is not None and self.reducer.world_size > 1:
            self.reducer.bucket_ready(i)

    def zero_grad(self, set_to_none: bool = False) -> None:   # gradients are views of the flat buffer: never set to None
        self.flat.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        from . import train_ops as T
        loss = closure() if closure is not None else None
        self._check_homes()
        scale = self.reducer.finish()
        g = self.param_groups[0]
        self.step_count += 1
        if not (self.lr_dev.is_cuda and torch.cuda.is_current_stream_capturing()):
            self._sync_lr()                      # under capture the fill would be baked into the graph: CapturedTrainStep syncs before replays
        T.adam_step(self.flat.flat_param, self.flat.flat_grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0], g["betas"][1],
                    g["eps"], g["weight_decay"], self.step_count, scale, step_dev=self.step_dev, lr_dev=self.lr_dev)
        self._invalidate_shadows()
        return loss

    # ---- torch.optim.Adam checkpoint format (hpe/main_h36m_lifting.py:239-242 loads torch.load(path)["optimizer"] into the optimizer)
    def state_dict(self):
        """Same layout as ``torch.optim.Adam.state_dict()``: per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` (copies of the flat
        moment buffers), so checkpoints written here resume under the reference's optimizer and vice versa."""
        step = int(self.step_dev.item())
        state = {}
        for i, p in enumerate(self.flat.params):
            o, n = self.flat.offsets[id(p)]
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        groups[0]["params"] = list(range(len(self.flat.params)))
        return {"state": state if step > 0 else {}, "param_groups": groups}

    def load_state_dict(self, state_dict) -> None:
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.flat.params):
            raise ValueError("optimizer state does not match this model (one parameter group over all parameters expected)")
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for i, idx in enumerate(groups[0]["params"]):
            st = state_dict["state"].get(idx)
            if st is None:
                continue
            p = self.flat.params[i]
            o, n = self.flat.offsets[id(p)]
            self.exp_avg[o:o + n].view(p.shape).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + n].view(p.shape).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}); one fused step count is kept")
        self.step_count = steps.pop() if steps else 0
        self.step_dev.fill_(self.step_count)
        self.broadcast_state()

    def _invalidate_shadows(self) -> None:
        """The kernel wrote the parameters behind autograd's back (their ``_version`` did not move): drop the cached 16-bit weight
        shadows / stacked head tensors so that the next forward rebuilds them."""
        for m in self._module.modules():
            if hasattr(m, "_shadow_key"):
                m._shadow_key = None
            if hasattr(m, "_head_key"):
                m._head_key = None
            if hasattr(m, "_fold_key"):
                m._fold_key = None


class CapturedTrainStep:
    """One training step — zero_grad, forward, objective, backward, (single-GPU) Adam — captured once as a CUDA graph and replayed.

    The step is shape-static (fixed batch and clip length) and every kernel launch, allocation and the optimizer's step counter live
    on the device, so a replay is one ``cudaGraphLaunch`` instead of ~570 launches of host work (10.2 ms vs 13.7 ms per BASELINE
    config 4 step on one B200).  Inputs are copied into static buffers before each replay; ``loss`` is a device scalar.

    Data parallel (world size > 1): the bucketed NCCL all-reduces of ``GradientReducer`` are captured with the step (NCCL kernels are
    graph nodes on NCCL's stream, forked from / joined to the capture by the events ``all_reduce(async_op=True)`` / ``wait()`` record), so
    every rank replays the same sequence of collectives and the host cost of the 18 NCCL launches disappears with the rest.  The
    capture is thread-local because the NCCL watchdog thread polls its own events meanwhile."""

    def __init__(self, model: nn.Module, optimizer: "FusedAdam", loss_fn, x: torch.Tensor, y: torch.Tensor, warmup: int = 3):
        self.x, self.y = x.clone(), y.clone()
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        distributed = optimizer.reducer.world_size > 1
        # The warm-up steps are real optimizer steps on the example batch.  They must not leak into training (a resumed checkpoint
        # would otherwise have been trained `warmup` extra steps on one batch): parameters, moments, the step count and the RNG
        # stream (stochastic depth) are put back once the graph exists.
        o = optimizer
        saved = (o.flat.flat_param.clone(), o.exp_avg.clone(), o.exp_avg_sq.clone(), o.step_dev.clone(), o.step_count,
                 torch.cuda.get_rng_state(self.x.device))
        for _ in range(max(1, warmup)):          # allocates shadows, tensor maps, workspaces (and NCCL communicators) outside the capture
            self._step()
        torch.cuda.synchronize()
        if distributed:
            dist.barrier(group=optimizer.reducer.group)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local" if distributed else "global"):
            self._step()
        torch.cuda.synchronize()
        o.flat.flat_param.copy_(saved[0])
        o.exp_avg.copy_(saved[1])
        o.exp_avg_sq.copy_(saved[2])
        o.step_dev.copy_(saved[3])
        o.step_count = saved[4]
        torch.cuda.set_rng_state(saved[5], self.x.device)
        o._invalidate_shadows()

    def _step(self) -> None:
        self.optimizer.zero_grad()
        out = self.model(self.x)
        loss = self.loss_fn(out, self.y)
        loss.backward()
        self.optimizer.step()
        self.loss = loss.detach()

    def close(self) -> None:
        """Destroys the graph.  Data parallel: call it BEFORE ``dist.destroy_process_group()`` — NCCL does not tear a communicator down
        while a graph that captured its collectives is alive (the destroy call blocks)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if self.graph is None:
            raise RuntimeError("CapturedTrainStep was closed")
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.optimizer._sync_lr()                # a scheduler may have moved param_groups[0]["lr"] since the last replay
        self.graph.replay()
        self.optimizer.step_count += 1           # host mirror of step_dev (state_dict reads the device value)
        return self.loss
