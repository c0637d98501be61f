"""Host -> device staging for the hot path's inputs.

The path's inputs (encoder_out, predictor_out, lengths, labels) are tens of MB per step; copied on
the compute stream they cost as much as the kernels.  ``HostBatchPrefetcher`` moves batch i+1 on
its own CUDA stream while batch i is being computed (pinned host memory, persistent device slots,
no allocation and no host synchronisation per step), which is how ``bench.py`` measures the
end-to-end number.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, List, Optional

import torch


class HostBatchPrefetcher:
    """Ring of ``depth`` device slots.

    ``put(host_batch)`` enqueues the copies into the next slot on a private stream (after the work that
    last read that slot); ``get()`` makes the current stream wait for the oldest enqueued batch and
    returns its device tensors.  The tensors returned by the previous ``get()`` are considered free
    once the next ``get()`` is called: everything that reads them has been enqueued by then.
    """

    def __init__(self, device: torch.device, depth: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("HostBatchPrefetcher needs a CUDA device (no CPU fallback on this path)")
        assert depth >= 2
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._free = [None] * depth  # event after which the slot may be overwritten
        self._next = 0
        self._queue: deque = deque()
        self._in_use: Optional[int] = None

    def _slot_for(self, idx: int, host_batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        slot = self._slots[idx]
        if slot is None or any(k not in slot or slot[k].shape != v.shape or slot[k].dtype != v.dtype
                               for k, v in host_batch.items()):
            # allocated ON the copy stream (the stream that writes them first): a block the caching allocator hands
            # back from the compute stream could still have queued readers there; get() records the compute stream
            # as a user, so a later reallocation (shapes change every step for real speech batches) waits for it
            with torch.cuda.stream(self.copy_stream):
                slot = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host_batch.items()}
            self._slots[idx] = slot
        return slot

    def put(self, host_batch: Dict[str, torch.Tensor]) -> None:
        if len(self._queue) + (1 if self._in_use is not None else 0) >= len(self._slots):
            raise RuntimeError("all device slots are busy: call get() before the next put()")
        for k, v in host_batch.items():
            if not v.is_pinned():
                raise ValueError(f"host tensor '{k}' must live in pinned memory for an asynchronous copy")
        idx = self._next
        self._next = (idx + 1) % len(self._slots)
        slot = self._slot_for(idx, host_batch)
        if self._free[idx] is not None:
            self.copy_stream.wait_event(self._free[idx])
        with torch.cuda.stream(self.copy_stream):
            for k, v in host_batch.items():
                slot[k].copy_(v, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._queue.append((idx, done))

    def get(self) -> Dict[str, torch.Tensor]:
        cur = torch.cuda.current_stream(self.device)
        if self._in_use is not None:  # the previous batch's readers are all enqueued on `cur` by now
            ev = torch.cuda.Event()
            ev.record(cur)
            self._free[self._in_use] = ev
        idx, done = self._queue.popleft()
        cur.wait_event(done)
        self._in_use = idx
        for t in self._slots[idx].values():
            t.record_stream(cur)
        return self._slots[idx]

    def __len__(self) -> int:
        return len(self._queue)
