"""Data-parallel glue for training on several GPUs of one box (the reference is single-device).

One process per GPU (torchrun), every rank holds the full model and a disjoint slice of the batch of videos.  The forward has
no cross-video operation except train-mode BatchNorm statistics (kept per replica, like DistributedDataParallel without
SyncBatchNorm), so the only exchange is the gradient:

* ``tggcn_backward_ex`` writes all parameter gradients into ONE flat buffer laid out in the order the backward completes them
  (four buckets: heads + segment level, frame level, embeddings, geometry GCN) and records an event per bucket;
* ``GradientAllReduce`` is called back as soon as the backward has been queued and issues one NCCL all-reduce per bucket on a
  side stream that waits for that bucket's event, so the reduction of the segment-level gradients (45 % of the bytes) runs
  under the frame-level backward instead of after it; ``reduce()`` only joins the streams;
* ``loss_term_weights`` makes the result the SINGLE-PROCESS gradient of the global batch: every loss term of the reference is a
  mean over the valid (target != -1) elements of the LOCAL batch (pyrutils/torch/losses.py:13-20, :47), so rank r's term i is
  scaled by n_i^r * world / sum_r n_i^r before the backward and the gradients are averaged.

Parameters off the gradient path (the 22-24 dead tensors of the reference) are not part of the flat
buffer on any rank: the layout is a pure function of the constructor arguments and of which segmentations are passed.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

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


def valid_counts(targets: Sequence[torch.Tensor], ignore_value: float = -1.0) -> torch.Tensor:
    """Per loss term, the number of target elements that take part in its mean (``target != -1``: the mask of
    pyrutils/torch/losses.py:13,30 and ``ignore_index=-1`` of :47).  Float64 tensor on the targets' device."""
    return torch.stack([(t != ignore_value).sum() for t in targets]).to(torch.float64)


def loss_term_weights(targets: Sequence[torch.Tensor], group=None, ignore_value: float = -1.0) -> torch.Tensor:
    """Factor for each local loss term so that the AVERAGE over ranks of the weighted terms (and of their gradients) equals the
    single-process loss over the global batch: w_i^r = n_i^r * world / sum_r n_i^r (0 where no rank has a valid element).
    One all-reduce of len(targets) doubles; issue it right after the targets are on the device so it overlaps the forward."""
    n = valid_counts(targets, ignore_value)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return torch.ones_like(n, dtype=torch.float32)
    total = n.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return torch.where(total > 0, n * world / total.clamp(min=1.0), torch.zeros_like(n)).to(torch.float32)


class GradientAllReduce:
    """Averages the model's flat gradient buffer over the process group, bucket by bucket, overlapped with the backward.

    ``attach()`` registers the reducer as the model's ``grad_ready_callback``; between ``loss.backward()`` and
    ``optimizer.step()`` call ``reduce()``.  ``extra_parameters``: tensors outside the model that also train (the learnable loss
    weights of a multi-task-loss module, pyrutils/torch/multi_task.py) — all-reduced as one small flat tensor in ``reduce()``.
    Without ``attach()`` (or with ``overlap=False``) ``reduce()`` falls back to one all-reduce of the whole buffer."""

    def __init__(self, model, group: Optional[dist.ProcessGroup] = None, overlap: bool = True,
                 extra_parameters: Iterable[torch.nn.Parameter] = ()):
        self.model, self.group, self.overlap = model, group, overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.extra = [p for p in extra_parameters]
        self._works: List = []
        self._launched_for = None
        self._comm_stream = None
        backend = dist.get_backend(group) if dist.is_initialized() else None
        self._avg = dist.ReduceOp.AVG if backend == 'nccl' else None          # gloo has no AVG: SUM then divide

    def attach(self):
        self.model.grad_ready_callback = self._on_backward_queued
        return self

    @torch.no_grad()
    def sync_parameters(self, src: int = 0):
        """Make every replica start from rank ``src``'s parameters and buffers (one broadcast per tensor, once)."""
        if self.world == 1:
            return
        for t in list(self.model.parameters()) + list(self.model.buffers()) + self.extra:
            dist.broadcast(t.data, src=src, group=self.group)

    def _all_reduce(self, t: torch.Tensor, async_op: bool):
        if self._avg is not None:
            return dist.all_reduce(t, op=self._avg, group=self.group, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=False)
        t.div_(self.world)
        return w

    @torch.no_grad()
    def _on_backward_queued(self, model):
        """Called by TGGCN._backward right after tggcn_backward_ex returned (everything is queued, nothing has to have run)."""
        if self.world == 1 or not self.overlap:
            return
        flat = model.flat_grad
        if not flat.is_cuda or self._avg is None:
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=flat.device)
        self._works = []
        for start, end, event in model.grad_buckets:
            self._comm_stream.wait_event(event)                        # bucket complete on the compute stream
            with torch.cuda.stream(self._comm_stream):
                self._works.append(self._all_reduce(flat[start:end], async_op=True))
        flat.record_stream(self._comm_stream)
        self._launched_for = flat.data_ptr()

    @torch.no_grad()
    def reduce(self):
        """Call between ``loss.backward()`` and ``optimizer.step()``.  Returns the (averaged) flat gradient buffer."""
        flat = self.model.flat_grad
        if flat is None:
            raise RuntimeError('no backward has run: model.flat_grad is empty')
        if self.world > 1:
            if self._works and self._launched_for == flat.data_ptr():
                for w in self._works:
                    w.wait()                                            # the current stream waits for the NCCL stream
            else:
                self._all_reduce(flat, async_op=False)
            self._works, self._launched_for = [], None
            if self.extra:
                grads = [p.grad for p in self.extra if p.grad is not None]
                if grads:
                    buf = torch.cat([g.reshape(-1) for g in grads])
                    self._all_reduce(buf, async_op=False)
                    off = 0
                    for g in grads:
                        g.copy_(buf[off:off + g.numel()].view_as(g))
                        off += g.numel()
        return flat
