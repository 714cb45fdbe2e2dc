"""Fused criterion: drop-in for ``vhoi.losses.select_loss`` / ``decide_num_main_losses`` (vhoi/losses.py:8-70, :103-112).

``select_loss('2G-GCN', model_input_type, dataset_name, cfg)`` returns ``(criterion, loss_names)`` with the reference's weights
and names; ``criterion(output, target, reduction='mean')`` returns the same list of weighted losses as
``pyrutils.torch.losses.multi_task_loss`` (budget, masked BCE, NLL with ignore_index -1), but computed by two kernels for all
outputs at once (``tggcn_loss_fwd`` / ``tggcn_loss_bwd``) and without the reference's ``.item()`` host synchronisations.
The returned losses carry autograd history, so ``sum(losses).backward()`` (pyrutils/torch/train_utils.py:150-151) works and
feeds the model's hand-written backward.  CUDA tensors only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import abi

BUDGET, BCE, NLL = 0, 1, 2


class LossTerm(C.Structure):
    _fields_ = [('kind', C.c_int32), ('weight', C.c_float), ('out', C.c_void_p), ('target', C.c_void_p), ('d_out', C.c_void_p),
                ('numel', C.c_int64), ('B', C.c_int32), ('C', C.c_int32), ('T', C.c_int32), ('E', C.c_int32)]


def _bind():
    lib = abi.lib()
    if not getattr(lib, '_tggcn_loss_bound', False):
        lib.tggcn_loss_fwd.restype = C.c_int
        lib.tggcn_loss_fwd.argtypes = [C.POINTER(LossTerm), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.tggcn_loss_bwd.restype = C.c_int
        lib.tggcn_loss_bwd.argtypes = [C.POINTER(LossTerm), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib._tggcn_loss_bound = True
    return lib


def _terms(kinds, weights, outputs, targets, d_outs):
    arr = (LossTerm * len(kinds))()
    for i, (k, w, o, t) in enumerate(zip(kinds, weights, outputs, targets)):
        arr[i].kind, arr[i].weight = k, float(w)
        arr[i].out, arr[i].target = o.data_ptr(), t.data_ptr()
        arr[i].d_out = d_outs[i].data_ptr() if d_outs is not None and d_outs[i] is not None else None
        if k == NLL:
            B, Cn, T, E = o.shape
            arr[i].numel, arr[i].B, arr[i].C, arr[i].T, arr[i].E = B * T * E, B, Cn, T, E
        else:
            arr[i].numel = o.numel()
    return arr


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kinds, weights, targets, *outputs):
        lib = _bind()
        dev = outputs[0].device
        outs = [o if (o.dtype == torch.float32 and o.is_contiguous()) else o.contiguous().float() for o in outputs]
        tgts = []
        for k, t in zip(kinds, targets):
            want = torch.int64 if k == NLL else torch.float32
            tgts.append(t if (t.dtype == want and t.is_contiguous()) else t.contiguous().to(want))
        losses = torch.empty(len(kinds), dtype=torch.float32, device=dev)
        scratch = torch.empty(2 * len(kinds), dtype=torch.float32, device=dev)
        terms = _terms(kinds, weights, outs, tgts, None)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            abi.check(lib.tggcn_loss_fwd(terms, len(kinds), losses.data_ptr(), scratch.data_ptr(), stream), 'tggcn_loss_fwd')
        ctx.kinds, ctx.weights, ctx.outs, ctx.tgts, ctx.scratch = kinds, weights, outs, tgts, scratch
        ctx.needs = [o.requires_grad for o in outputs]
        return losses

    @staticmethod
    def backward(ctx, g):
        lib = _bind()
        dev = g.device
        g = g.contiguous().float()
        d_outs = [torch.empty_like(o) if need else None for o, need in zip(ctx.outs, ctx.needs)]
        terms = _terms(ctx.kinds, ctx.weights, ctx.outs, ctx.tgts, d_outs)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            abi.check(lib.tggcn_loss_bwd(terms, len(ctx.kinds), ctx.scratch.data_ptr(), g.data_ptr(), stream), 'tggcn_loss_bwd')
        return (None, None, None) + tuple(d_outs)


def multi_task_loss(input: Sequence[torch.Tensor], target: Sequence[torch.Tensor], kinds: Sequence[int],
                    weight: Sequence[float] = None, reduction: str = 'mean') -> List[torch.Tensor]:
    """Same contract as pyrutils.torch.losses.multi_task_loss for the loss tuple of the 2G-GCN model."""
    if reduction != 'mean':
        raise NotImplementedError("only reduction='mean' (what train_utils.py:147 passes) is supported")
    if not input[0].is_cuda:
        raise abi.TggcnError('the fused criterion runs on CUDA tensors only (no CPU fallback)')
    n = min(len(input), len(target), len(kinds))
    weight = [1.0] * n if weight is None else list(weight)[:n]
    losses = _FusedLoss.apply(tuple(kinds[:n]), tuple(weight), tuple(target[:n]), *input[:n])
    return list(losses.unbind(0))


class _Criterion:
    def __init__(self, kinds, weight):
        self.kinds, self.weight = list(kinds), list(weight)

    def __call__(self, output, target, reduction='mean'):
        return multi_task_loss(output, target, self.kinds, self.weight, reduction)


def _get(cfg, key, default):
    """cfg may be an omegaconf-1.4 node (``get(k, default_value=)``), a dict, or None."""
    if cfg is None:
        return default
    try:
        return cfg.get(key, default_value=default)
    except TypeError:
        return cfg.get(key, default)


def select_loss(model_name: str, model_input_type: str, dataset_name: str, cfg):
    """vhoi/losses.py:8-70 for the 2G-GCN model: same weights from ``cfg.misc``, same loss names."""
    if model_name != '2G-GCN':
        raise ValueError(f'Unknown model {model_name}: only the 2G-GCN criterion is provided by the B200 path')
    misc = _get(cfg, 'misc', {}) or {}
    sub = lambda group, key, default: (_get(misc, group, {}) or {}).get(key, default)
    cad = dataset_name == 'cad120'
    hb = ob = 0.0
    if sub('budget_loss', 'add', False):
        hb, ob = sub('budget_loss', 'human_weight', 1.0), sub('budget_loss', 'object_weight', 1.0)
    weight = [hb, ob] if cad else [hb]
    hs = os_ = 0.0
    s_weight = sub('segmentation_loss', 'weight', 1.0)
    add_seg = sub('segmentation_loss', 'add', False)
    if add_seg and not _get(misc, 'input_human_segmentation', False):
        hs = s_weight
    if add_seg and not _get(misc, 'input_object_segmentation', False):
        os_ = s_weight
    weight += [hs, os_] if cad else [hs]
    weight_val = 0.0 if (add_seg and sub('segmentation_loss', 'pretrain', False)) else 1.0
    ant = _get(misc, 'anticipation_loss_weight', 1.0)
    fl = _get(misc, 'first_level_loss_weight', 0.0)
    if cad:
        weight += [fl] * 4 + [weight_val, ant, weight_val, ant]
        kinds = [BUDGET, BUDGET, BCE, BCE] + [NLL] * 8
        names = ['B_HS', 'B_OS', 'BCE_HS', 'BCE_OS', 'NLL_SAR_F', 'NLL_SAP_F', 'NLL_OAR_F', 'NLL_OAP_F',
                 'NLL_SAR', 'NLL_SAP', 'NLL_OAR', 'NLL_OAP']
    else:
        weight += [fl] * 2 + [weight_val, ant]
        kinds = [BUDGET, BCE] + [NLL] * 4
        names = ['B_HS', 'BCE_HS', 'NLL_SAR_F', 'NLL_SAP_F', 'NLL_SAR', 'NLL_SAP']
    return _Criterion(kinds, weight), names


def decide_num_main_losses(model_name: str, dataset_name: str, misc_dict: dict):
    """vhoi/losses.py:103-112."""
    if model_name != '2G-GCN':
        return None
    seg = (misc_dict or {}).get('segmentation_loss', {}) or {}
    if seg.get('add', False) and seg.get('pretrain', False):
        return 10 if dataset_name == 'cad120' else 5
    return 4 if dataset_name == 'cad120' else 2
