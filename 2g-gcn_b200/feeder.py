"""Host->device input path.

The reference moves every batch with a synchronous ``.to(device)`` from pageable memory right before the forward
(vhoi/data_loading.py:1282-1315, ``pin_memory=False``).  At GPU speed that copy (51 MB per MPHOI batch, ~1 ms over PCIe) is
15 % of a step.  ``DeviceBatchPipeline`` keeps two device-side buffer sets and copies batch i+1 from pinned host memory on a
side stream while batch i computes; events order the two streams in both directions (copy -> compute before the forward reads
a slot, compute -> copy before the slot is overwritten).  The tensors handed to the model are ordinary device tensors, so the
model and the unchanged feeder (``gcn_forward``) do not know about it.
"""
from __future__ import annotations

from typing import Dict, List

import torch


class DeviceBatchPipeline:
    def __init__(self, device, example: Dict[str, torch.Tensor], depth: int = 2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots: List[Dict[str, torch.Tensor]] = [
            {k: torch.empty_like(v, device=self.device) for k, v in example.items()} for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]      # copy finished  (copy stream -> compute stream)
        self.free = [torch.cuda.Event() for _ in range(depth)]       # consumer done  (compute stream -> copy stream)
        self._submitted = 0
        self._taken = 0
        self._has_free = [False] * depth
        self.bytes_per_batch = sum(v.numel() * v.element_size() for v in example.values())

    def submit(self, host_batch: Dict[str, torch.Tensor]) -> None:
        """Start copying ``host_batch`` (pinned tensors) into the next slot; returns immediately."""
        if self._submitted - self._taken >= self.depth:
            raise RuntimeError('DeviceBatchPipeline: all slots are in flight; call get()/release() first')
        j = self._submitted % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self._has_free[j]:
                self.copy_stream.wait_event(self.free[j])            # the forward that read this slot has run
            for k, dst in self.slots[j].items():
                dst.copy_(host_batch[k], non_blocking=True)
            self.ready[j].record(self.copy_stream)
        self._submitted += 1

    def get(self) -> Dict[str, torch.Tensor]:
        """Device tensors of the oldest submitted batch; the current stream waits for its copy."""
        if self._taken >= self._submitted:
            raise RuntimeError('DeviceBatchPipeline: nothing submitted')
        j = self._taken % self.depth
        torch.cuda.current_stream(self.device).wait_event(self.ready[j])
        self._current = j
        self._taken += 1
        return self.slots[j]

    def release(self) -> None:
        """Call after the work that reads the batch returned by the last get() has been enqueued on the current stream."""
        j = self._current
        self.free[j].record(torch.cuda.current_stream(self.device))
        self._has_free[j] = True
