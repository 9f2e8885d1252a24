"""Data parallelism for the D-MPNN hot path: one process per GPU, each rank collates
and processes its OWN molecules (shard before collation — a molecule's result depends on its local
batch because of the reference's node-offset ``rev_index``), and the only exchange per training step
is the all-reduce of the flat fp32 gradient buffer (NCCL over NVLink 5 / NVSwitch on GPUs; gloo in
the CPU tests). Inference needs no collective.

The buffer is cut into *buckets* (contiguous slices, one per parameter group the caller names). With
``overlap=True`` a bucket's all-reduce is issued from a post-accumulate-grad hook the moment its last
gradient has been written — for the message-passing block that is right after layer 0's weight
gradient (K4b), i.e. while layer 0's dgrad / backward epilogue and the embedding backward still run —
on NCCL's own stream; ``finish()`` joins it before the optimizer. The mean is taken by the collective
itself (``ReduceOp.AVG``) on NCCL, so there is no separate divide kernel; gloo has no AVG and gets
SUM followed by an in-place divide.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

__all__ = ["FlatGradients", "shard_range", "shard_list"]


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous ``[lo, hi)`` slice of ``n`` molecules owned by ``rank`` (sizes differ by at most 1)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} out of range for world size {world_size}")
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def shard_list(items: Sequence, rank: int, world_size: int) -> Sequence:
    lo, hi = shard_range(len(items), rank, world_size)
    return items[lo:hi]


class FlatGradients:
    """All parameter gradients as views into ONE flat buffer; the data-parallel exchange is one all-reduce per bucket
    (1.08 MB at d=300 L=3, 21 MB at d=1024 L=5: latency-, not bandwidth-bound).

    ``params``: an iterable of parameters (one bucket) or a list of iterables (one bucket each, laid out in that order).
    ``overlap``: issue each bucket's all-reduce from autograd hooks as soon as the bucket is complete (see module docstring);
    call ``finish()`` after ``backward()``. Without it, ``all_reduce_mean()`` reduces everything at once."""

    def __init__(self, params: Iterable, process_group=None, overlap: bool = False):
        params = list(params)
        groups = [list(g) for g in params] if params and not isinstance(params[0], Tensor) else [params]
        seen: set[int] = set()
        self.buckets: list[list[Tensor]] = []
        for g in groups:
            uniq = []
            for p in g:
                if p.requires_grad and id(p) not in seen:  # a shared layer appears once
                    seen.add(id(p))
                    uniq.append(p)
            if uniq:
                self.buckets.append(uniq)
        self.params = [p for b in self.buckets for p in b]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("all parameters must share one device and dtype")
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=dt, device=dev)
        self.slices: list[tuple[int, int]] = []
        off = 0
        for b in self.buckets:
            lo = off
            for p in b:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.slices.append((lo, off))
        self.group = process_group
        self.overlap = bool(overlap)
        self._pending: list = []
        self._fired = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._hooks = []
        if self.overlap:
            for bi, b in enumerate(self.buckets):
                for p in b:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    # ---- bookkeeping ------------------------------------------------------------------------
    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self) -> None:
        """Use this (or ``zero_grad(set_to_none=False)``) instead of the default ``zero_grad()``: ``set_to_none=True`` drops the
        views and autograd would then allocate gradients OUTSIDE the flat buffer."""
        self.flat.zero_()

    def rebind(self) -> int:
        """Re-attach every ``p.grad`` to its slice of the flat buffer. A gradient that autograd allocated elsewhere (after a
        ``zero_grad(set_to_none=True)``) is copied in first. Returns the number of gradients that had to be re-bound."""
        moved, off = 0, 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                else:
                    view.zero_()
                p.grad = view
                moved += 1
        return moved

    def _check_bound(self, bucket: int | None = None) -> None:
        off = 0
        first = 0 if bucket is None else sum(len(b) for b in self.buckets[:bucket])
        last = len(self.params) if bucket is None else first + len(self.buckets[bucket])
        for i, p in enumerate(self.params):
            if not first <= i < last:
                off += p.numel()
                continue
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                raise RuntimeError(
                    f"FlatGradients: the gradient of parameter {i} {tuple(p.shape)} no longer aliases the flat buffer (zero_grad(set_to_none=True)?); "
                    "all-reducing it would exchange stale zeros. Use flat.zero() / zero_grad(set_to_none=False), or call flat.rebind().")
            off += p.numel()

    # ---- the exchange -----------------------------------------------------------------------
    def _has_avg(self) -> bool:
        try:
            return dist.get_backend(self.group) == "nccl"
        except Exception:
            return False

    def _reduce(self, t: Tensor, async_op: bool):
        w = self.world_size
        if self._has_avg():
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op), None
        work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return work, (t, w)

    def _make_hook(self, bi: int):
        def hook(_param):
            self._fired[bi] += 1
            if self._fired[bi] == len(self.buckets[bi]) and not self._launched[bi]:
                self._launch_bucket(bi)
        return hook

    def _launch_bucket(self, bi: int) -> None:
        self._check_bound(bi)  # raises out of backward() when a gradient was re-allocated outside the buffer
        self._launched[bi] = True
        if self.world_size == 1:
            return
        lo, hi = self.slices[bi]
        self._pending.append(self._reduce(self.flat[lo:hi], async_op=True))

    def finish(self) -> None:
        """Overlap mode: reduce the buckets whose hooks did not all fire (unused parameters), then make the current stream wait
        for every bucket. Resets the per-step state."""
        if not self.overlap:
            self.all_reduce_mean()
            return
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:
                self._launch_bucket(bi)
        for work, post in self._pending:
            if work is not None:
                work.wait()
            if post is not None:
                post[0].div_(post[1])
        self.reset()

    def reset(self) -> None:
        """Forget the per-step hook state (after an exception inside backward)."""
        self._pending.clear()
        self._fired = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

    def all_reduce_mean(self, async_op: bool = False):
        """Mean over ranks of the whole buffer in one collective (what DDP does). No-op for one rank."""
        self._check_bound()
        w = self.world_size
        if w == 1:
            return None
        work, post = self._reduce(self.flat, async_op=async_op)
        if async_op:
            return _Averaged(work, post)
        if post is not None:
            post[0].div_(post[1])
        return None


class _Averaged:
    def __init__(self, work, post):
        self.work, self.post = work, post

    def wait(self) -> None:
        self.work.wait()
        if self.post is not None:
            self.post[0].div_(self.post[1])
