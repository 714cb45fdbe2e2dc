"""Data-parallel glue for training on several GPUs of one box (the reference is single-device).

One process per GPU (torchrun), every rank holds the full model and a disjoint slice of the batch of videos.  The
forward has no cross-video operation except train-mode BatchNorm statistics (kept per replica, like
torch.nn.parallel.DistributedDataParallel does without SyncBatchNorm), so the only exchange is the gradient:
``tggcn_backward`` writes all parameter gradients into ONE flat buffer (``model.flat_grad``), which is all-reduced with
a single NCCL call over NVLink and rebound to the parameters' ``.grad`` before ``optimizer.step()``.

Parameters that are off the gradient path (the 22-24 dead tensors of the reference) are not part
of the flat buffer on any rank, so there is nothing to skip "consistently": the layout is a pure function of the
constructor arguments and of which segmentations are passed.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Contiguous slice of the videos for this rank.  Shard AFTER padding to the global max length: the padded length
    changes the result through the un-permuted view of the geometry features (vhoi/models.py:644-645)."""
    out = {}
    for k, v in batch.items():
        if not torch.is_tensor(v) or v.dim() == 0:
            out[k] = v
            continue
        n = v.size(0)
        if n % world != 0:
            raise ValueError(f'{k}: batch of {n} videos does not split over {world} ranks')
        per = n // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


class GradientAllReduce:
    """Averages ``model.flat_grad`` over the process group after each backward and rebinds the parameter gradients."""

    def __init__(self, model, group: Optional[dist.ProcessGroup] = None):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    @torch.no_grad()
    def sync_parameters(self, src: int = 0):
        """Make every replica start from rank ``src``'s parameters and buffers (one broadcast per tensor, once)."""
        if self.world == 1:
            return
        for t in list(self.model.parameters()) + list(self.model.buffers()):
            dist.broadcast(t.data, src=src, group=self.group)

    @torch.no_grad()
    def reduce(self):
        """Call between ``loss.backward()`` and ``optimizer.step()``."""
        flat = self.model.flat_grad
        if flat is None:
            raise RuntimeError('no backward has run: model.flat_grad is empty')
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
        self.model.bind_flat_grads()
        return flat
