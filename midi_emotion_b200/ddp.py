"""Plain data-parallel training for the model (one process per GPU, NCCL allreduce over NVLink).

The reference is single-device (train.py:33-34); BASELINE.json asks for DDP only.  Gradients are
averaged across ranks in flat buckets.  The hot path has no other exchange step.
"""
from __future__ import annotations

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
    in a hooked buffer (e.g. when gradients were accumulated into pre-existing `.grad` tensors)."""

    def __init__(self, model: torch.nn.Module, group=None, bucket_mb: float = 256.0, overlap: bool = True):
        self.module = model
        self.group = group
        self.bucket_mb = bucket_mb
        self._pending = []
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1:
            with torch.no_grad():
                for p in model.parameters():
                    dist.broadcast(p, src=0, group=group)
            if overlap and hasattr(model, "_param_list") and next(model.parameters()).is_cuda:
                model._grad_ready_hook = self._on_grads_ready

    def __call__(self, *a, **kw):
        return self.module(*a, **kw)

    def _on_grads_ready(self, flat: torch.Tensor) -> None:
        work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        self._pending.append((flat, work))

    def sync_gradients(self) -> None:
        if self.world == 1:
            return
        done = set()
        for flat, work in self._pending:
            work.wait()
            done.add(flat.untyped_storage().data_ptr())
        self._pending = []
        rest = [p.grad for p in self.module.parameters()
                if p.grad is not None and p.grad.untyped_storage().data_ptr() not in done]
        allreduce_mean_(rest, self.group, self.bucket_mb)


def shard_batch(n_items: int, rank: int, world: int):
    """Even contiguous split of a global batch (DistributedSampler-style, drop_last)."""
    per = n_items // world
    return range(rank * per, (rank + 1) * per)
