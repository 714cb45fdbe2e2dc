"""Synthetic MPHOI-72 / CAD-120 / Bimanual-shaped batches and deterministic weights.

The datasets are not in the reference tree, so every measurement and parity test runs on synthetic
tensors with the layout ``vhoi/data_loading.py`` produces:

* ``x_human``  (B, T, H, 2048 + 4V): pooled ROI feature ‖ geometry tail (V nodes × [x, y, vx, vy],
  identical for every human of a frame, data_loading.py:836-839), frames past a video's length are
  zero (NaN padding zeroed on host, data_loading.py:373, :849);
* ``x_objects`` (B, T, O, 2048), ``objects_mask`` (B, O), ``steps_per_example`` (B,);
* targets: segmentation target (B, T, H) in [0,1] with -1 padding, int64 labels (B, T, E) with -1 padding.

Everything is generated on the CPU generator so the same seed gives the same bytes here and on the
GPU box (same torch build).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch


@dataclass(frozen=True)
class Shape:
    name: str
    H: int
    O: int
    V: int
    num_classes: Tuple[int, Optional[int]]
    hh: bool            # message_humans_to_human
    dataset: str

    @property
    def Fh(self) -> int:
        return 2048 + 4 * self.V


MPHOI = Shape('mphoi', 2, 4, 26, (13, None), True, 'mphoi')
CAD120 = Shape('cad120', 1, 5, 19, (10, 12), False, 'cad120')
BIMANUAL = Shape('bimanual', 2, 9, 30, (14, None), True, 'bimanual')
SHAPES = {s.name: s for s in (MPHOI, CAD120, BIMANUAL)}


def model_kwargs(shape: Shape, hidden_size: int = 512, stage: int = 1, **overrides) -> dict:
    """Constructor kwargs == ``cfg.parameters`` of conf/models/2G-GCN_stage{1,2}.yaml:4-29 plus the
    ``input_size`` / ``num_classes`` that train.py:28-34 adds."""
    kw = dict(
        input_size=(shape.Fh, 2048), num_classes=shape.num_classes,
        add_segment_length=0, add_time_position=0, time_position_strategy='s', positional_encoding_style='e',
        attention_style='v3', bias=True, cat_level_states=0, discrete_networks_num_layers=1,
        discrete_optimization_strategy='gs', filter_discrete_updates=(stage == 2), gcn_node=shape.V,
        hidden_size=hidden_size, message_humans_to_human=shape.hh, message_human_to_objects=True,
        message_objects_to_human=True, message_objects_to_object=True, message_geometry_to_objects=True,
        message_geometry_to_human=False, message_segment=True, message_type='v2', message_granularity='v1',
        message_aggregation='att', object_segment_update_strategy='ind', share_level_mlps=0,
        update_segment_threshold=0.1 if stage == 2 else 0.5)
    overrides = {k: v for k, v in overrides.items() if not k.startswith('_')}      # '_…' keys are test-case options, not kwargs
    kw.update(overrides)          # e.g. cat_level_states=1, share_level_mlps=1 (yaml: conf/models/2G-GCN_stage1.yaml:12,28)
    return kw


def make_batch(shape: Shape, B: int, T: int, seed: int = 1234, min_len_frac: float = 0.6,
               min_objects: int = 2) -> Dict[str, torch.Tensor]:
    """One padded batch with the statistics SURVEY.md §8d prescribes."""
    g = torch.Generator().manual_seed(seed)
    H, O, V = shape.H, shape.O, shape.V
    lo = max(1, math.ceil(min_len_frac * T))
    lengths = torch.randint(lo, T + 1, (B,), generator=g)
    lengths[0] = T
    n_obj = torch.randint(min(min_objects, O), O + 1, (B,), generator=g)
    vis = torch.relu(torch.randn(B, T, H, 2048, generator=g))
    pos = torch.rand(B, T, V, 2, generator=g) * 1.5
    vel = torch.randn(B, T, V, 2, generator=g)
    xo = torch.relu(torch.randn(B, T, O, 2048, generator=g))
    om = torch.zeros(B, O)
    for b in range(B):
        L = int(lengths[b])
        vel[b, L - 1] = 0.0                                  # last real frame has zero velocity
        vis[b, L:] = 0.0
        pos[b, L:] = 0.0
        vel[b, L:] = 0.0
        xo[b, L:] = 0.0
        om[b, :int(n_obj[b])] = 1.0
        xo[b, :, int(n_obj[b]):] = 0.0
    geo = torch.cat([pos, vel], dim=-1).reshape(B, T, 1, 4 * V).expand(B, T, H, 4 * V)
    xh = torch.cat([vis, geo], dim=-1).contiguous()
    return dict(x_human=xh, x_objects=xo.contiguous(), objects_mask=om,
                steps_per_example=lengths.to(torch.float32), lengths=lengths)


def make_distances(shape: Shape, B: int, T: int, seed: int = 77):
    """Synthetic entity distances for misc.make_attention_distance_based (vhoi/data_loading.py:1264-1276): (hh, ho, oo) with
    shapes (B,T,H,H) [None for CAD-120, whose loader has none], (B,T,H,O), (B,T,O,O); a tenth of the entries are exactly 0
    ('no distance available': masked by compute_distance_based_attention_weights, vhoi/models.py:1769-1772)."""
    g = torch.Generator().manual_seed(seed)
    H, O = shape.H, shape.O

    def one(*size):
        d = 0.05 + torch.rand(*size, generator=g)
        return torch.where(torch.rand(*size, generator=g) < 0.1, torch.zeros_like(d), d)
    hh = one(B, T, H, H) if shape.dataset != 'cad120' else None
    return hh, one(B, T, H, O), one(B, T, O, O)


def make_targets(shape: Shape, lengths: torch.Tensor, T: int, seed: int = 4321) -> Dict[str, torch.Tensor]:
    """Piecewise-constant labels (segments of 8-40 frames) and the Gaussian-smoothed (sigma=4)
    end-frame indicator of data_loading.py:545-559; -1 past each video's length."""
    g = torch.Generator().manual_seed(seed)
    B = lengths.numel()

    def labels(E: int, C: int):
        y = torch.full((B, T, E), -1, dtype=torch.int64)
        ends = torch.zeros(B, T, E)
        for b in range(B):
            L = int(lengths[b])
            for e in range(E):
                t = 0
                while t < L:
                    seg = int(torch.randint(8, 41, (1,), generator=g))
                    c = int(torch.randint(0, C, (1,), generator=g))
                    y[b, t:min(t + seg, L), e] = c
                    t += seg
                    ends[b, min(t, L) - 1, e] = 1.0
        return y, ends

    def smooth(ends: torch.Tensor, sigma: float = 4.0) -> torch.Tensor:
        r = int(4.0 * sigma + 0.5)
        k = torch.exp(-0.5 * (torch.arange(-r, r + 1, dtype=torch.float32) / sigma) ** 2)
        k = k / k.sum()
        x = ends.permute(0, 2, 1).reshape(-1, 1, T)
        s = torch.nn.functional.conv1d(x, k.view(1, 1, -1), padding=r)
        s = (s * 2.5 * sigma).clamp(0.0, 1.0).reshape(B, -1, T).permute(0, 2, 1).contiguous()
        for b in range(B):
            s[b, int(lengths[b]):] = -1.0
        return s

    C_h, C_o = shape.num_classes
    y_h, ends_h = labels(shape.H, C_h)
    y_h_pred = torch.roll(y_h, -1, dims=1)
    for b in range(B):
        y_h_pred[b, int(lengths[b]) - 1:] = y_h[b, int(lengths[b]) - 1:]
    out = dict(seg_h=smooth(ends_h), rec_h=y_h, pred_h=y_h_pred)
    if C_o is not None:
        y_o, ends_o = labels(shape.O, C_o)
        out.update(seg_o=smooth(ends_o), rec_o=y_o, pred_o=y_o.clone())
    return out


def target_list(shape: Shape, tg: Dict[str, torch.Tensor]):
    """Targets in the positional order ``multi_task_loss`` zips with the model outputs
    (data_loading.py:517-519; vhoi/losses.py:41-60)."""
    if shape.num_classes[1] is None:
        return [tg['seg_h'], tg['seg_h'], tg['rec_h'], tg['pred_h'], tg['rec_h'], tg['pred_h']]
    return [tg['seg_h'], tg['seg_o'], tg['seg_h'], tg['seg_o'],
            tg['rec_h'], tg['pred_h'], tg['rec_o'], tg['pred_o'],
            tg['rec_h'], tg['pred_h'], tg['rec_o'], tg['pred_o']]


def deterministic_fill(state_dict: Dict[str, torch.Tensor], seed: int = 0, gain: float = 1.0) -> None:
    """Overwrite every entry of a (reference-layout) ``state_dict`` in place with values that depend
    only on (seed, key, shape) — not on module construction order — so the reference model, the oracle
    and the CUDA model can be given identical weights on any machine.
    Linear/conv/GRU weights: U(-a, a), a = gain/sqrt(fan_in); biases U(-.05, .05)*gain; BatchNorm: weight U(.5,1.5), bias/mean
    U(-.5,.5), var U(.5,1.5)."""
    with torch.no_grad():
        for idx, key in enumerate(sorted(state_dict.keys())):
            v = state_dict[key]
            g = torch.Generator().manual_seed(seed * 1000003 + idx)
            if key.endswith('num_batches_tracked'):
                v.zero_()
                continue
            u = torch.rand(v.shape, generator=g, dtype=torch.float32)
            if '.bn.' in key:
                if key.endswith('weight') or key.endswith('running_var'):
                    val = 0.5 + u
                else:
                    val = u - 0.5
            else:
                if key.endswith('bias') or '.bias_' in key:
                    a = 0.05 * gain
                else:
                    fan_in = v[0].numel() if v.dim() > 1 else v.numel()
                    if key == 'geometry_embedding_gcn.weight':
                        fan_in = v.shape[1]                     # models_gcn.py:26-28 uses size(1)
                    a = gain / math.sqrt(fan_in)
                val = (2.0 * u - 1.0) * a
            v.copy_(val.to(v.dtype))



def state_checksum(state_dict: Dict[str, torch.Tensor]) -> float:
    """A cheap order-independent fingerprint used to check that weights regenerated on another box are
    the same bytes as the ones the golden vectors were made with."""
    acc = 0.0
    for key in sorted(state_dict.keys()):
        v = state_dict[key].detach().to('cpu', torch.float64).reshape(-1)
        if v.numel():
            w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 97 + 1.0
            acc += float((v * w).sum())
    return acc
