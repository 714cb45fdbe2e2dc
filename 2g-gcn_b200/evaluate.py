"""Evaluation post-processing on the device.

``predict.py`` brings every output to the host, up-samples it with ``torch.repeat_interleave`` to the target's frame rate
(predict.py:64-70, ``match_shape`` :95-122), takes ``np.argmax`` over the classes (``process_output`` :186-202) and scores the
per-frame labels with the segmental F1@k of ``pyrutils/metrics.py:7-81`` on (video x entity) rows (predict.py:229-246).  For
large sweeps the same three steps run here as two kernels on the tensors the model just produced; only the per-row scores
(one double per row and overlap) travel to the host, where they are averaged in row order like the reference does.

There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence

import torch

from . import abi


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def predict_labels(output: torch.Tensor, target: torch.Tensor, downsampling: int = 1) -> torch.Tensor:
    """(B,C,T,E) log-probabilities -> (B,Tt,E) int64 labels at the target's frame rate (Tt = target.size(1)):
    ``np.argmax(match_shape(repeat_interleave(out, downsampling, dim=-2), tgt), axis=1)``."""
    if not output.is_cuda:
        raise abi.TggcnError('evaluate.predict_labels runs on a CUDA device only')
    if output.ndim != 4:
        raise RuntimeError(f'Number of dimensions for output is {output.ndim}')      # predict.py:66
    B, Cn, T, E = output.shape
    Tt = int(target.size(1))
    out = output.detach().contiguous().float()
    labels = torch.empty(B, Tt, E, dtype=torch.int64, device=output.device)
    with torch.cuda.device(output.device):
        abi.check(abi.lib().tggcn_upsample_argmax(out.data_ptr(), labels.data_ptr(), B, Cn, T, E, Tt, int(downsampling),
                                                  _stream(output.device)), 'tggcn_upsample_argmax')
    return labels


def f1_at_k(target: torch.Tensor, pred: torch.Tensor, num_classes: int, overlaps: Sequence[float] = (0.10, 0.25, 0.50),
            ignore_value: int = -1) -> Dict[float, float]:
    """F1@k of (B,Tt,E) int64 label tensors for every overlap in ``overlaps`` — ``pyrutils.metrics.f1_at_k`` on the rows
    ``np.swapaxes(x, 1, 2).reshape(-1, Tt)`` with ``ignore_value=-1.0`` (predict.py:229-246)."""
    if not (target.is_cuda and pred.is_cuda):
        raise abi.TggcnError('evaluate.f1_at_k runs on a CUDA device only')
    if target.shape != pred.shape or target.ndim != 3:
        raise ValueError(f'target {tuple(target.shape)} and prediction {tuple(pred.shape)} must both be (B, T, E)')
    dev = target.device
    B, Tt, E = target.shape
    tgt = target.contiguous().long()
    prd = pred.contiguous().long()
    ov = torch.tensor(list(overlaps), dtype=torch.float64, device=dev)
    rows = B * E
    scratch = torch.empty(int(abi.lib().tggcn_f1_at_k_scratch_bytes(B, Tt, E)), dtype=torch.uint8, device=dev)
    f1 = torch.empty(len(overlaps), rows, dtype=torch.float64, device=dev)
    valid = torch.empty(rows, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        abi.check(abi.lib().tggcn_f1_at_k(tgt.data_ptr(), prd.data_ptr(), B, Tt, E, int(num_classes), ov.data_ptr(), len(overlaps),
                                          int(ignore_value), scratch.data_ptr(), f1.data_ptr(), valid.data_ptr(), _stream(dev)),
                  'tggcn_f1_at_k')
    f1_host, valid_host = f1.cpu(), valid.cpu().bool()
    n = int(valid_host.sum())
    result = {}
    for k, ovl in enumerate(overlaps):
        total = 0.0
        for v in f1_host[k][valid_host].tolist():          # sequential sum in row order, as metrics.py:72-81 accumulates
            total += v
        result[float(ovl)] = total / n                     # ZeroDivisionError when nothing is valid, like the reference
    return result


def evaluate_outputs(outputs: Sequence[torch.Tensor], targets: Sequence[torch.Tensor], num_classes: Sequence[int],
                     downsampling: int = 1, overlaps: Sequence[float] = (0.10, 0.25, 0.50)):
    """predict.py's scoring of the main outputs: per output i, labels at the target's frame rate, frame accuracy over the
    valid frames (the micro-averaged F1 of ``evaluate_predictions`` :205-226) and F1@k.  Returns a list of dicts."""
    results = []
    for out, tgt, nc in zip(outputs, targets, num_classes):
        labels = predict_labels(out, tgt, downsampling)
        keep = tgt != -1
        acc = float((labels[keep] == tgt[keep]).double().mean()) if bool(keep.any()) else float('nan')
        results.append({'labels': labels, 'accuracy': acc, 'f1_at_k': f1_at_k(tgt, labels, nc, overlaps)})
    return results
