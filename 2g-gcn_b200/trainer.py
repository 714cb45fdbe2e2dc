"""Data-parallel training driver: what ``pyrutils.torch.train_utils.train`` (train_utils.py:12-115) does
for one device, for N GPUs of one box — one process per GPU under torchrun.

Kept from the reference so that ``train.py``'s surroundings stay valid:
  * the caller's objects: ``fetch_model_data`` / ``feed_model_data`` (vhoi.data_loading.select_model_data_fetcher / _feeder),
    ``criterion`` returning a LIST of losses that is summed (train_utils.py:148-150), ``mtll_model``, ``num_main_losses``,
    ``clip_gradient_at``;
  * the returned checkpoint dictionary: ``epoch``, ``model_state_dict``, ``mtll_model_state_dict``, ``train_losses``,
    ``val_losses``, ``train_raw_losses``, ``val_raw_losses`` with the reference's "lowest validation loss wins" rule
    (train_utils.py:94-112), so ``torch.save(checkpoint, ...)`` files interchange with the reference's.

Different, because a B200 runs a step in milliseconds:
  * the dataset lives on the GPU (feeder.DeviceResidentDataset); a seeded ``ShardedBatchSampler`` hands every rank its slice of
    each GLOBAL batch (videos keep the dataset-wide padded length);
  * gradients are averaged by dp.GradientAllReduce — bucket all-reduces launched from the backward's stage boundaries, overlapping
    the rest of the backward — after every rank's loss terms were weighted by its share of the valid target elements
    (dp.loss_term_weights), which makes the update equal to the single-process update on the global batch;
  * nothing on the step path synchronises: losses are accumulated on the device, log lines are printed from values copied to
    pinned memory asynchronously (they appear one log interval late), the status words of the persistent kernels are checked at
    every log line and at the end of each epoch;
  * the two extra full ``test()`` passes per epoch of the reference (train_utils.py:58, :81) become one sharded evaluation pass
    whose per-term sums are all-reduced once (``eval_train_set=False`` drops the pass over the training set, whose running means
    are already known from the training steps).
"""
from __future__ import annotations

import math
import os
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import dp as dp_mod
from .feeder import DeviceResidentDataset


class ShardedBatchSampler:
    """Index lists for one rank.  Every epoch draws one permutation of the videos from ``seed + epoch`` (identical on all ranks),
    cuts it into global batches of ``global_batch`` videos and gives rank r the r-th contiguous ``global_batch / world`` slice of
    each.  The tail of an epoch is filled up by wrapping around the permutation, so every rank runs the same number of steps
    with the same local batch size (a collective per step needs that); ``len()`` = steps per epoch."""

    def __init__(self, num_videos: int, global_batch: int, rank: int = 0, world: int = 1, shuffle: bool = True, seed: int = 0):
        if global_batch % world != 0:
            raise ValueError(f'global batch {global_batch} does not split over {world} ranks')
        if num_videos < 1:
            raise ValueError('empty dataset')
        self.n, self.global_batch, self.rank, self.world = num_videos, global_batch, rank, world
        self.local_batch = global_batch // world
        self.shuffle, self.seed = shuffle, seed
        self.epoch = 0

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def __len__(self):
        return math.ceil(self.n / self.global_batch)

    def __iter__(self):
        if self.shuffle:
            order = torch.randperm(self.n, generator=torch.Generator().manual_seed(self.seed + self.epoch)).tolist()
        else:
            order = list(range(self.n))
        steps = len(self)
        need = steps * self.global_batch
        order = (order * math.ceil(need / self.n))[:need]
        for s in range(steps):
            lo = s * self.global_batch + self.rank * self.local_batch
            yield order[lo:lo + self.local_batch]


class _AsyncScalars:
    """Device scalars -> pinned host memory without a synchronisation; ``ready()`` yields what has landed."""

    def __init__(self):
        self.pending = []

    def push(self, values: torch.Tensor, meta):
        if values.is_cuda:
            host = torch.empty(values.shape, dtype=values.dtype).pin_memory()
            host.copy_(values, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(values.device))
        else:
            host, ev = values.clone(), None
        self.pending.append((host, ev, meta))

    def ready(self, wait: bool = False):
        out, keep = [], []
        for host, ev, meta in self.pending:
            if ev is not None and wait:
                ev.synchronize()
            if ev is None or ev.query():
                out.append((host, meta))
            else:
                keep.append((host, ev, meta))
        self.pending = keep
        return out


class DataParallelTrainer:
    def __init__(self, model, optimizer, criterion: Callable, loss_names: Sequence[str], device,
                 fetch_model_data: Callable, feed_model_data: Callable, mtll_model=None, clip_gradient_at: float = 0.0,
                 num_main_losses: Optional[int] = None, group=None, log_interval: int = 25, overlap: bool = True,
                 verbose: bool = True, feed_kwargs: Optional[dict] = None):
        self.model, self.optimizer, self.criterion = model, optimizer, criterion
        self.loss_names, self.device = list(loss_names), torch.device(device)
        self.fetch, self.feed, self.mtll = fetch_model_data, feed_model_data, mtll_model
        self.clip, self.num_main, self.group = clip_gradient_at, num_main_losses, group
        self.log_interval, self.verbose = log_interval, verbose
        self.feed_kwargs = dict(feed_kwargs or {})
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        extra = list(mtll_model.parameters()) if mtll_model is not None else []
        self.reducer = dp_mod.GradientAllReduce(model, group=group, overlap=overlap, extra_parameters=extra).attach()
        self.reducer.sync_parameters()
        self._log = _AsyncScalars()

    # ---- data ----------------------------------------------------------------------------------------------------------------
    def stage(self, dataset) -> DeviceResidentDataset:
        """``TensorDataset`` (or a sequence of tensors) -> device-resident copy.  Every rank stages the whole dataset: it is small
        next to HBM and lets the per-epoch permutation pick any video for any rank without host traffic."""
        if isinstance(dataset, DeviceResidentDataset):
            return dataset
        tensors = dataset.tensors if hasattr(dataset, 'tensors') else dataset
        return DeviceResidentDataset(tensors, self.device)

    def _say(self, *a, **k):
        if self.verbose and self.rank == 0:
            print(*a, **k)

    def _check_status(self, wait: bool):
        """Status words of the persistent kernels: without ``wait`` only the calls whose words have already landed are looked at
        (no synchronisation); with ``wait`` everything queued so far (end of an epoch / of an evaluation pass)."""
        if getattr(self.model, '_last', None) is None:
            return
        if wait:
            self.model.check_persistent_kernels()
        else:
            self.model._poll_status()

    # ---- one epoch of training -----------------------------------------------------------------------------------------------
    def train_epoch(self, data: DeviceResidentDataset, sampler: ShardedBatchSampler) -> torch.Tensor:
        """Returns the per-term running means of the (globally weighted) losses over the epoch, a device tensor."""
        self.model.train()
        if self.mtll is not None:
            self.mtll.train()
        steps = len(sampler)
        sums = None
        for step, indices in enumerate(sampler):
            data_in, target = self.fetch(data.batch(indices), device=self.device)
            weights = dp_mod.loss_term_weights(target, group=self.group)        # tiny all-reduce, overlaps the forward
            self.optimizer.zero_grad(set_to_none=True)
            output = self.feed(self.model, data_in, **self.feed_kwargs)
            losses = self.criterion(output, target, reduction='mean')
            losses = [l * w for l, w in zip(losses, weights.to(losses[0].device).unbind(0))]
            if self.mtll is not None:
                losses = self.mtll(losses)
            total = sum(losses)
            total.backward()                     # TGGCN._backward queues the backward, then the bucket all-reduces
            self.reducer.reduce()
            if self.clip:
                params = list(self.model.parameters()) + (list(self.mtll.parameters()) if self.mtll is not None else [])
                torch.nn.utils.clip_grad_norm_(params, max_norm=self.clip)
            self.optimizer.step()
            vec = torch.stack([l.detach() for l in losses])
            sums = vec if sums is None else sums + vec
            if step % self.log_interval == 0 or step == steps - 1:
                self._log.push(vec, (step, steps, sampler))
                self._print_ready()
                self._check_status(wait=False)
        self._print_ready(wait=True)
        self._check_status(wait=True)
        return sums / steps

    def _print_ready(self, wait: bool = False):
        for host, (step, steps, sampler) in self._log.ready(wait):
            vals = host          # this rank's weighted terms; the exact global values come with the epoch summary
            n_main = self.num_main if self.num_main is not None else len(vals)
            seen = min((step + 1) * sampler.global_batch, sampler.n)
            line = f'(Train) Batch [{seen:6d}/{sampler.n:6d} ({100 * (step + 1) / steps:3.0f}%)]  Loss: {float(vals[-n_main:].sum()): 8.4f}'
            line += ''.join(f'  {name}: {float(v): 6.4f}' for name, v in zip(self.loss_names, vals))
            self._say(line)

    # ---- evaluation (the reference's test(), train_utils.py:168-227), sharded -------------------------------------------------
    @torch.no_grad()
    def evaluate(self, data: DeviceResidentDataset, global_batch: int, name: str = 'Test'):
        """Mean of the per-batch losses like the reference (test() divides the summed batch losses by the number of batches),
        with the batches dealt round-robin to the ranks.  Returns (total, per-term list, raw total, raw per-term list); the raw
        values are None without an mtll model."""
        self.model.eval()
        if self.mtll is not None:
            self.mtll.eval()
        sampler = ShardedBatchSampler(len(data), global_batch, 0, 1, shuffle=False)
        sums = raw_sums = None
        count = 0
        for b, indices in enumerate(sampler):
            if b % self.world != self.rank:
                continue
            indices = indices[:max(1, min(len(indices), len(data) - b * global_batch))]      # the last batch is not wrapped here
            data_in, target = self.fetch(data.batch(indices), device=self.device)
            output = self.feed(self.model, data_in, **self.feed_kwargs)
            raw = self.criterion(output, target, reduction='mean')
            losses = self.mtll(raw) if self.mtll is not None else raw
            vec = torch.stack([l.detach() for l in losses]).double()
            sums = vec if sums is None else sums + vec
            if self.mtll is not None:
                rvec = torch.stack([l.detach() for l in raw]).double()
                raw_sums = rvec if raw_sums is None else raw_sums + rvec
            count += 1
        n_terms = len(self.loss_names)
        pack = torch.zeros(2 * n_terms + 1, dtype=torch.float64, device=self.device)
        if sums is not None:
            pack[:n_terms] = sums
            pack[-1] = count
        if raw_sums is not None:
            pack[n_terms:2 * n_terms] = raw_sums
        if self.world > 1:
            dist.all_reduce(pack, group=self.group)
        pack = pack.cpu()                              # the one synchronisation of the pass
        self._check_status(wait=True)
        n_batches = max(float(pack[-1]), 1.0)
        terms = [float(v) / n_batches for v in pack[:n_terms]]
        n_main = self.num_main if self.num_main is not None else n_terms
        total = sum(terms[-n_main:])
        tag = f'({name})'
        self._say(f'{tag:>12} Loss: {total: 7.4f}' + ''.join(f'   {n}: {v: 6.4f}' for n, v in zip(self.loss_names, terms)))
        if self.mtll is None:
            return total, terms, None, None
        raw_terms = [float(v) / n_batches for v in pack[n_terms:2 * n_terms]]
        return total, terms, sum(raw_terms[-n_main:]), raw_terms

    # ---- the epoch loop ---------------------------------------------------------------------------------------------------------
    def fit(self, train_dataset, epochs: int, global_batch: int, val_dataset=None, initial_epoch: int = 1, seed: int = 0,
            eval_train_set: bool = True, checkpoint_path: Optional[str] = None) -> dict:
        train_data = self.stage(train_dataset)
        val_data = self.stage(val_dataset) if val_dataset is not None else None
        sampler = ShardedBatchSampler(len(train_data), global_batch, self.rank, self.world, shuffle=True, seed=seed)
        ckpt = {}
        history = {'train_losses': [], 'val_losses': [], 'train_raw_losses': [], 'val_raw_losses': []}
        best_val = float('inf')
        last = initial_epoch + epochs - 1
        for epoch in range(initial_epoch, initial_epoch + epochs):
            self._say(f'\nEpoch: [{epoch:4d}/{last:4d}]')
            sampler.set_epoch(epoch)
            running = self.train_epoch(train_data, sampler)
            if eval_train_set:
                tot, terms, raw_tot, raw_terms = self.evaluate(train_data, global_batch, 'Train')
            else:                                   # running means of the training steps, averaged over the ranks
                r = running.double()
                if self.world > 1:
                    dist.all_reduce(r, group=self.group)
                    r /= self.world
                terms = [float(v) for v in r.cpu()]
                n_main = self.num_main if self.num_main is not None else len(terms)
                tot, raw_tot, raw_terms = sum(terms[-n_main:]), None, None
            history['train_losses'].append([tot, terms])
            if self.mtll is not None and raw_terms is not None:
                history['train_raw_losses'].append([raw_tot, raw_terms])
            take = val_data is None
            if val_data is not None:
                vtot, vterms, vraw_tot, vraw_terms = self.evaluate(val_data, global_batch, 'Validation')
                history['val_losses'].append([vtot, vterms])
                if self.mtll is not None:
                    history['val_raw_losses'].append([vraw_tot, vraw_terms])
                if vtot < best_val:
                    best_val, take = vtot, True
            if take:                                # snapshot on the host: later steps update the parameters in place
                ckpt['epoch'] = epoch
                ckpt['model_state_dict'] = {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}
                if self.mtll is not None:
                    ckpt['mtll_model_state_dict'] = {k: v.detach().cpu().clone() for k, v in self.mtll.state_dict().items()}
        self._say('Lowest val_loss is', best_val)
        ckpt.update(history)
        if checkpoint_path is not None and self.rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(checkpoint_path)), exist_ok=True)
            torch.save(ckpt, checkpoint_path)
        if self.world > 1:
            dist.barrier(group=self.group)
        return ckpt
