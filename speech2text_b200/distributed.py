"""Multi-GPU plumbing for the transducer-loss path: one process per GPU, utterances
sharded across ranks, no data-path collective.  The only exchanges are the ones the
reference gets from Lightning DDP (rnnt_task.py:506-512):
one all-reduce of the joiner weight gradients (kept in ONE flat buffer so it is a
single NCCL call over NVLink, no bucketing copies) and one of the scalar losses.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of ``range(n_items)`` for ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


class FlatGradBucket:
    """All gradients of ``params`` live in one flat fp32 buffer; ``p.grad`` are views into it.

    autograd accumulates into an existing ``.grad`` in place, so after ``zero()`` a backward pass
    leaves the flat buffer holding [dW_enc, db_enc, dW_pre, db_pre, dW1, db1, dW2, db2] ready for
    a single all-reduce.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], extra_scalars: int = 0):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        # ``extra_scalars`` slots at the tail carry the logged losses (train_loss, simple_loss, pruned_loss:
        # rnnt_task.py:506-512, sync_dist=True) through the SAME all-reduce as the gradients
        self.n_grad = total
        self.flat = torch.zeros(total + extra_scalars, dtype=torch.float32, device=dev)
        self.scalars = self.flat[total:]
        self.bound = False
        self.written = set()  # id(param) of the bound sinks a backward pass has overwritten since the last zero()
        self.attach()

    def attach(self) -> None:
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            if getattr(self, "bound", False):
                p._s2t_grad_sink = (p.grad, self)
            off += n

    def bind(self) -> "FlatGradBucket":
        """Let the path's backward kernels write each gradient straight into its slice of the flat buffer
        (the dW kernels write into the buffer NCCL sends, no extra copy).  A bound parameter's
        gradient is OVERWRITTEN by every backward pass of the tensor-core path and autograd's accumulate is
        skipped for it (eight read-modify-write passes per step less); parameters that reach autograd through plain
        torch ops (strict-fp32 mode projections) still accumulate, so ``zero()`` stays a real zero.

        The overwrite only happens when it equals what autograd would have produced
        (``functional.claim_grad_sinks``): the first backward pass after ``zero()`` while ``p.grad`` still
        aliases the buffer.  A second backward pass before the next ``zero()`` (gradient accumulation,
        ``retain_graph``) and a parameter whose ``.grad`` was reset (``zero_grad(set_to_none=True)``) fall back to
        autograd's accumulate, so bound and unbound buckets give the same gradients; call ``attach()`` to
        re-alias the parameters after such a reset."""
        for p in self.params:
            p._s2t_grad_sink = (p.grad, self)
        self.bound = True
        return self

    def unbind(self) -> None:
        for p in self.params:
            if hasattr(p, "_s2t_grad_sink"):
                del p._s2t_grad_sink
        self.bound = False

    def zero(self) -> None:
        self.flat.zero_()
        self.written.clear()

    def put_scalars(self, values: Sequence[torch.Tensor]) -> torch.Tensor:
        """Store a few 0-d tensors in the tail slots (one small kernel): they are averaged over the ranks by the
        gradient all-reduce itself instead of by a second collective."""
        assert len(values) == self.scalars.numel(), (len(values), self.scalars.numel())
        torch.stack([v.detach().float().reshape(()) for v in values], out=self.scalars)
        return self.scalars

    def all_reduce(self, average: bool = True, group=None, async_op: bool = False):
        """ONE collective for the whole step: gradients and logged scalars, averaged inside NCCL (``ReduceOp.AVG``:
        no separate division pass over the buffer).  Capturable in a CUDA graph together with the step."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if average and dist.get_backend(group) == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        if average:  # gloo has no AVG
            self.flat.div_(dist.get_world_size(group))
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def reduce_scalars(values: Sequence[torch.Tensor], group=None) -> torch.Tensor:
    """Mean over ranks of a few 0-d tensors in ONE all-reduce (the reference logs
    train_loss, simple_loss, pruned_loss with sync_dist=True)."""
    vec = torch.stack([v.detach().float().reshape(()) for v in values])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        vec /= dist.get_world_size(group)
    return vec
