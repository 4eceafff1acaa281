"""Plain data-parallel training for the model (one process per GPU, NCCL allreduce over NVLink).

The reference is single-device (train.py:33-34); BASELINE.json asks for DDP only.  Gradients are
averaged across ranks in flat buckets.  The hot path has no other exchange step.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def _buckets(tensors: List[torch.Tensor], cap_bytes: int) -> List[List[torch.Tensor]]:
    out, cur, size = [], [], 0
    for t in tensors:
        n = t.numel() * t.element_size()
        if cur and size + n > cap_bytes:
            out.append(cur)
            cur, size = [], 0
        cur.append(t)
        size += n
    if cur:
        out.append(cur)
    return out


def allreduce_mean_(tensors: Iterable[torch.Tensor], group=None, bucket_mb: float = 256.0) -> None:
    """In-place average of `tensors` over the process group, reduced in flat buckets (reverse order,
    i.e. the order backward produced them)."""
    tensors = [t for t in tensors if t is not None]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    for bucket in _buckets(list(reversed(tensors)), int(bucket_mb * (1 << 20))):
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n


class DataParallel:
    """Minimal DDP wrapper: identical initial weights (broadcast from rank 0), gradient averaging
    overlapped with backward, `sync_gradients()` after `loss.backward()` and before clipping / the
    optimiser step (train.py:319-325 order).

    The model's hand-written backward produces the gradients of one layer in one flat buffer and calls
    `model._grad_ready_hook(flat)` as soon as they are enqueued; the hook starts an asynchronous NCCL
    allreduce (AVG) of that buffer, which runs while the earlier layers' backward kernels execute.
    `sync_gradients()` waits for those reductions and reduces, in buckets, whatever gradient did not live
    in a hooked buffer (e.g. when gradients were accumulated into pre-existing `.grad` tensors).

    Gradient accumulation (the reference's `accumulate_step > 1`, train.py:314-319): wrap the non-boundary
    micro-steps in `with ddp.no_sync():` -- nothing is reduced there -- and call `sync_gradients()` once before the
    optimiser step; the accumulated `.grad` tensors are then reduced post hoc.  A backward pass that finds `.grad`
    already populated never starts an overlapped reduction (AccumulateGrad's `p.grad += g` would race with it),
    and first makes the compute stream wait for reductions an earlier backward pass of the same step started;
    since averaging is linear and idempotent on already-averaged values, avg(avg(g1) + g2) = avg(g1) + avg(g2)."""

    def __init__(self, model: torch.nn.Module, group=None, bucket_mb: float = 256.0, overlap=None):
        # overlap: True = start each layer's allreduce from the backward pass; "defer" = collect the layers' flat
        # gradient buffers and reduce them in place, back to back, in sync_gradients(); False = no hook at all.
        # Default from ME_DDP_OVERLAP (1 / defer / 0), else True.
        if overlap is None:
            env = os.environ.get("ME_DDP_OVERLAP", "1")
            overlap = "defer" if env == "defer" else env != "0"
        self.module = model
        self.group = group
        self.bucket_mb = bucket_mb
        self.defer = overlap == "defer"
        self._deferred = []
        self._pending = []
        self.require_backward_grad_sync = True
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1:
            with torch.no_grad():
                for p in model.parameters():
                    dist.broadcast(p, src=0, group=group)
            if overlap and hasattr(model, "_param_list") and next(model.parameters()).is_cuda:
                model._grad_ready_hook = self._on_grads_ready

    def __call__(self, *a, **kw):
        return self.module(*a, **kw)

    def no_sync(self):
        """Context manager for the non-boundary micro-steps of gradient accumulation (torch DDP's name)."""
        ddp = self

        class _NoSync:
            def __enter__(self):
                self.prev = ddp.require_backward_grad_sync
                ddp.require_backward_grad_sync = False

            def __exit__(self, *exc):
                ddp.require_backward_grad_sync = self.prev
                return False

        return _NoSync()

    def _accumulating(self) -> bool:
        # the first parameter receives its gradient last (the input stage is the end of backward): a populated
        # .grad there means an earlier backward pass of this optimiser step has already run
        p0 = next(iter(self.module.parameters()))
        return p0.grad is not None

    def _on_grads_ready(self, flat: torch.Tensor) -> None:
        if not self.require_backward_grad_sync:
            return
        if self._accumulating():
            # `.grad += g` follows on the compute stream: finish what an earlier pass started, forget it (those
            # buffers now hold averaged + local parts and are reduced again, post hoc, by sync_gradients)
            self._drain()
            self._deferred = []
            return
        if self.defer:
            self._deferred.append(flat)
            return
        if dist.get_backend(self.group) == "nccl":
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._pending.append((flat, work, False))
        else:   # gloo (CPU tests) has no AVG: sum now, divide when the work is waited for
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((flat, work, True))

    def _drain(self):
        for flat in self._deferred:   # reduced in place now, one after the other (no compute runs beside them)
            if dist.get_backend(self.group) == "nccl":
                self._pending.append((flat, dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True), False))
            else:
                self._pending.append((flat, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), True))
        self._deferred = []
        done = set()
        for flat, work, divide in self._pending:
            work.wait()
            if divide:
                flat.div_(self.world)
            done.add(flat.untyped_storage().data_ptr())
        self._pending = []
        return done

    def sync_gradients(self) -> None:
        if self.world == 1:
            return
        done = self._drain()
        rest = [p.grad for p in self.module.parameters()
                if p.grad is not None and p.grad.untyped_storage().data_ptr() not in done]
        allreduce_mean_(rest, self.group, self.bucket_mb)


def shard_batch(n_items: int, rank: int, world: int):
    """Even contiguous split of a global batch (DistributedSampler-style, drop_last)."""
    per = n_items // world
    return range(rank * per, (rank + 1) * per)
