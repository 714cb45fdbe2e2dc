"""Data-parallel counterpart of ``train_single_epoch`` (pyrutils/torch/train_utils.py:118-165), SURVEY.md §8 row f1.

Same arguments and the same per-batch sequence as the reference (fetch -> zero_grad -> feed -> criterion -> sum -> backward ->
clip -> step -> log), with two differences that only matter at GPU speed:
  * ``reducer`` (``dp.GradientAllReduce``): between ``backward()`` and the optimiser, all-reduce the flat gradient buffer over
    the ranks of one box (one NCCL call) and rebind the parameter gradients;
  * the losses are only read back (``.item()``, a host synchronisation) when a line is printed, every ``log_interval``
    batches — the reference reads them on the same schedule, so the printed output is the same.
The unchanged reference loop keeps working with the drop-in model on one GPU; this one is for torchrun launches.
"""
from __future__ import annotations

import torch


def train_single_epoch(model, data_loader, optimizer, criterion, device, loss_names, clip_gradient_at=0.0,
                       fetch_model_data=None, feed_model_data=None, log_interval=25, mtll_model=None,
                       num_main_losses=None, reducer=None, verbose=True, **kwargs):
    """Trains ``model`` for one epoch; returns the list of summed per-batch losses (device tensors, no sync)."""
    if fetch_model_data is None or feed_model_data is None:
        raise ValueError('fetch_model_data and feed_model_data are required (vhoi.data_loading.select_model_data_fetcher/feeder)')
    model.train()
    if mtll_model is not None:
        mtll_model.train()
    num_examples = len(data_loader.dataset)
    history = []
    for batch_idx, dataset in enumerate(data_loader):
        data, target = fetch_model_data(dataset, device=device)
        optimizer.zero_grad()
        output = feed_model_data(model, data, **kwargs)
        losses = criterion(output, target, reduction='mean')
        if mtll_model is not None:
            losses = mtll_model(losses)
        loss = sum(losses)
        loss.backward()
        if reducer is not None:
            reducer.reduce()
        if clip_gradient_at:
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=clip_gradient_at)
        optimizer.step()
        history.append(loss.detach())
        log_now, is_last_batch = (batch_idx % log_interval) == 0, batch_idx == (len(data_loader) - 1)
        if verbose and (log_now or is_last_batch):
            n_main = num_main_losses if num_main_losses is not None else len(losses)
            main = sum(losses[-n_main:])
            first = min((batch_idx + 1) * data_loader.batch_size, num_examples)
            progress = 100 * (batch_idx + 1) / len(data_loader)
            print(f'(Train) Batch [{first:6d}/{num_examples:6d} ({progress:3.0f}%)] ', f'Loss: {main.item(): 8.4f}', end='')
            for loss_name, single_loss in zip(loss_names, losses):
                print(f'  {loss_name}: {single_loss: 6.4f}', end='')
            print()
    return history
