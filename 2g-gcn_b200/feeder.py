"""Host->device input path.

The reference materialises the whole dataset as fp32 tensors in pageable host memory (``TensorDataset``,
vhoi/data_loading.py:362-379, ``pin_memory=False``, ``num_workers=0``) and moves every batch with a synchronous ``.to(device)``
right before the forward (``gcn_fetcher``, :1282-1315).  At GPU speed that copy (51 MB per MPHOI batch, ~1 ms over PCIe) is
15 % of a step.  Two replacements, both handing ordinary device tensors to the unchanged fetcher / feeder:

``DeviceResidentDataset``  the datasets of the paper are small next to 180 GB of HBM (MPHOI-72: 72 videos, CAD-120: 120,
    Bimanual: 540; < 10 GB each after down-sampling), so the whole ``TensorDataset`` is staged ONCE through pinned memory and
    every batch is a device-side gather: zero host->device bytes per step.  Videos keep the dataset-wide padded length — the
    padded length changes results through the un-permuted geometry view (vhoi/models.py:644-645), so nothing is trimmed.

``DeviceBatchPipeline``  for data that does not fit (or arrives per step): ``depth`` device-side slots, batch i+1 is copied from
    pinned host memory on a side stream while batch i computes; events order the two streams in both directions.  Call
    ``submit`` for batch i+1 AFTER queueing the forward of batch i: the forward's own small host->device copy (the Gumbel draws)
    shares the copy engine and would wait behind the 51 MB batch (measured: 4.60 -> 5.23 ms per step the other way round).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch


class DeviceResidentDataset:
    """All tensors of a reference ``TensorDataset`` resident on ``device``; ``batch(indices)`` returns the tuple layout the
    unchanged ``gcn_fetcher`` indexes (vhoi/data_loading.py:517-519: 8 inputs then the targets)."""

    def __init__(self, tensors: Sequence[torch.Tensor], device, chunk_bytes: int = 256 << 20):
        self.device = torch.device(device)
        n = {int(t.size(0)) for t in tensors}
        if len(n) != 1:
            raise ValueError(f'all dataset tensors must have the same number of videos, got {sorted(n)}')
        self.num_videos = n.pop()
        self.tensors: List[torch.Tensor] = []
        self.staged_bytes = 0
        for t in tensors:
            dst = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            # stage through a bounded pinned buffer: pinning the whole dataset at once would double the host footprint
            rows = max(1, chunk_bytes // max(1, t[0].numel() * t.element_size())) if t.dim() > 0 and t.size(0) else 1
            for r0 in range(0, t.size(0), rows):
                src = t[r0:r0 + rows].contiguous()
                src = src.pin_memory() if self.device.type == 'cuda' and not src.is_pinned() else src
                dst[r0:r0 + rows].copy_(src, non_blocking=True)
                if self.device.type == 'cuda':
                    torch.cuda.current_stream(self.device).synchronize()      # the pinned chunk is reused / freed
            self.tensors.append(dst)
            self.staged_bytes += t.numel() * t.element_size()

    def __len__(self):
        return self.num_videos

    @classmethod
    def from_tensor_dataset(cls, dataset, device):
        return cls(dataset.tensors, device)

    def batch(self, indices) -> List[torch.Tensor]:
        idx = torch.as_tensor(indices, dtype=torch.int64, device=self.device)
        return [t.index_select(0, idx) for t in self.tensors]


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
        self._released = 0
        self.bytes_per_batch = sum(v.numel() * v.element_size() for v in example.values())

    def submit(self, host_batch: Dict[str, torch.Tensor]) -> None:
        """Start copying ``host_batch`` (pinned tensors) into the next slot; returns immediately.  A slot is free again only
        after ``release()``: a batch that was handed out by ``get()`` but not released still blocks its slot."""
        if self._submitted - self._released >= self.depth:
            raise RuntimeError('DeviceBatchPipeline: every slot holds a batch that has not been released; call get() and '
                               'release() (after the work that reads the batch has been queued) first')
        j = self._submitted % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self._submitted >= self.depth:
                self.copy_stream.wait_event(self.free[j])            # the work that read this slot's previous batch has run
            for k, dst in self.slots[j].items():
                dst.copy_(host_batch[k], non_blocking=True)
            self.ready[j].record(self.copy_stream)
        self._submitted += 1

    def get(self) -> Dict[str, torch.Tensor]:
        """Device tensors of the oldest submitted batch; the current stream waits for its copy."""
        if self._taken >= self._submitted:
            raise RuntimeError('DeviceBatchPipeline: nothing submitted')
        if self._taken != self._released:
            raise RuntimeError('DeviceBatchPipeline: release() the batch of the previous get() first')
        j = self._taken % self.depth
        torch.cuda.current_stream(self.device).wait_event(self.ready[j])
        self._taken += 1
        return self.slots[j]

    def release(self) -> None:
        """Call after ALL work that reads the batch of the last get() has been queued on the current stream.  In training that
        is after ``loss.backward()``: tggcn_backward reads x_human / x_objects straight from the slot (weight gradients of the
        embeddings), so releasing after the forward alone would let the next copy race with the backward."""
        if self._released >= self._taken:
            raise RuntimeError('DeviceBatchPipeline: release() without a matching get()')
        j = self._released % self.depth
        self.free[j].record(torch.cuda.current_stream(self.device))
        self._released += 1
