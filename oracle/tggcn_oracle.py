"""CPU oracle for the 2G-GCN (TGGCN) hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain-PyTorch (CPU, fp32/fp64) restatement of the algorithm of the reference's
``TGGCN.forward`` + losses + F1@k.  It is *not* part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and there only as the checker (or as the timed CPU baseline), never as the thing shipped.

Parity status: **pinned**.  The reference has no tests or golden vectors of its own, so
the pins are outputs of the *unmodified reference executed in the build container*
(``oracle/gen_golden.py`` imports ``/root/reference``) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against every one of them.

Every function cites the reference lines it restates (paths relative to the reference root).
The control flow deliberately keeps the reference's per-timestep structure (two Python loops over T),
so that timing this file on host cores is a fair stand-in for the reference's own CPU path.

State is a flat ``dict`` of tensors keyed by the reference's ``state_dict`` names.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from itertools import accumulate, groupby
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

_GCN = 'geometry_embedding_gcn.'


@dataclass
class OracleConfig:
    """The subset of ``conf/models/2G-GCN_stage{1,2}.yaml:4-29`` that changes the arithmetic."""
    hidden_size: int = 512
    gcn_node: int = 26
    num_classes: Tuple[int, Optional[int]] = (13, None)
    message_humans_to_human: bool = True
    filter_discrete_updates: bool = False
    update_segment_threshold: float = 0.5
    cat_level_states: bool = False          # models.py:901-903 (share_level_mlps changes no arithmetic: same tensors, two names)
    mean_pool: bool = False                 # message_aggregation 'mp' (models.py:1033-1036 and the other message functions)
    att_scaled: bool = True                 # attention_style 'v3' (scaled dot product); False = 'v2' (models.py:1740-1745)
    time_position: str = ''                 # add_time_position: '' off, 's' = time embedding appended to the segment-level inputs
                                            # (models.py:755-762), 'u' = appended to the gate MLP inputs (:656-662, :1494, :1527)
    positional_encoding: str = 'e'          # 'e' = time_position_mlp (Linear(1, D) + ReLU) of (t+1)/steps, 'p' = periodic of (t+1)
    segment_length: bool = False            # add_segment_length (models.py:763-779, :954-979): per entity, the (normalised) length of
                                            # the segment closed at a frame, embedded like the time position, appended to the xx rows
    gate_layers: int = 1                    # discrete_networks_num_layers (models.py:532-547): hidden Linear(in, D) + ReLU layers before
                                            # the Linear(., 1) + sigmoid of the gate MLPs
    geo_to_human: bool = False              # message_geometry_to_human (models.py:690-695, :1432-1475): one more message block
                                            # ReLU(geometry_to_human_message_mlp([x_g | h_g])) in the humans' segment inputs and gate inputs
    straight_through: bool = False          # discrete_optimization_strategy 'st' (models.py:1621-1622): soft gate = the sigmoid
                                            # probability itself, hard gate = (p > thr) with identity gradient, no Gumbel noise
    update_strategy: str = 'ind'            # object_segment_update_strategy 'ind' | 'sah' | 'coh' (models.py:1523-1532); 'sah' and
                                            # 'coh' only differ from 'ind' with exactly one human (models.py:741-742)


_UPD = {'independent': 'ind', 'ind': 'ind', 'same_as_human': 'sah', 'sah': 'sah', 'conditional_on_human': 'coh', 'coh': 'coh'}


def config_from_kwargs(kw: dict) -> OracleConfig:
    """OracleConfig of a TGGCN constructor-kwargs dict (``cfg.parameters`` of the yaml files + input_size / num_classes)."""
    return OracleConfig(kw['hidden_size'], kw['gcn_node'], tuple(kw['num_classes']), bool(kw['message_humans_to_human']),
                        bool(kw['filter_discrete_updates']), float(kw['update_segment_threshold']),
                        bool(kw.get('cat_level_states', 0)), kw.get('message_aggregation') in ('mp', 'mean_pooling'),
                        kw.get('attention_style') not in ('v2', 'dot-product'),
                        (kw.get('time_position_strategy', 's') if kw.get('add_time_position') else ''),
                        'e' if kw.get('positional_encoding_style', 'e') in ('e', 'embedding') else 'p',
                        bool(kw.get('add_segment_length', 0)),
                        int(kw.get('discrete_networks_num_layers', 1)),
                        bool(kw.get('message_geometry_to_human', False)),
                        kw.get('discrete_optimization_strategy', 'gs') in ('st', 'straight-through'),
                        _UPD[kw.get('object_segment_update_strategy', 'ind')])


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def _lin(p: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    """nn.Linear as built by build_mlp (pyrutils/torch/models.py:31-33): x W^T + b."""
    return x @ p[name + '.weight'].t() + p[name + '.bias']


def _relu_lin(p, name, x):
    return torch.relu(_lin(p, name, x))


def geo_gcn(p: Dict[str, Tensor], xg: Tensor, training: bool = False,
            stats_out: Optional[dict] = None) -> Tensor:
    """Geo_gcn.forward, pyrutils/torch/models_gcn.py:30-37 (+ norm_data :45-50, embed :57-62,
    compute_similarity :95-100).  ``xg`` is (B, 4, V, T); returns (B, 128, V, T) contiguous."""
    B, C, V, T = xg.shape
    # BatchNorm1d over channel index c*V+v, statistics over (B, T)  (models_gcn.py:46-48)
    x = xg.reshape(B, C * V, T)
    g, b = p[_GCN + 'joint_embed.cnn.0.bn.weight'], p[_GCN + 'joint_embed.cnn.0.bn.bias']
    if training:
        mean = x.mean(dim=(0, 2))
        var = x.var(dim=(0, 2), unbiased=False)
        if stats_out is not None:
            n = B * T
            stats_out['batch_mean'] = mean
            stats_out['batch_var_unbiased'] = var * n / max(n - 1, 1)
    else:
        mean = p[_GCN + 'joint_embed.cnn.0.bn.running_mean']
        var = p[_GCN + 'joint_embed.cnn.0.bn.running_var']
    x = (x - mean[None, :, None]) / torch.sqrt(var[None, :, None] + 1e-5) * g[None, :, None] + b[None, :, None]
    x = x.reshape(B, C, V, T).permute(0, 3, 2, 1)                     # (B, T, V, 4)
    w1 = p[_GCN + 'joint_embed.cnn.1.cnn.weight'].reshape(64, C)
    w3 = p[_GCN + 'joint_embed.cnn.3.cnn.weight'].reshape(64, 64)
    e = torch.relu(x @ w1.t() + p[_GCN + 'joint_embed.cnn.1.cnn.bias'])
    e = torch.relu(e @ w3.t() + p[_GCN + 'joint_embed.cnn.3.cnn.bias'])   # (B, T, V, 64)
    th = e @ p[_GCN + 'get_s.s1.cnn.weight'].reshape(128, 64).t() + p[_GCN + 'get_s.s1.cnn.bias']
    ph = e @ p[_GCN + 'get_s.s2.cnn.weight'].reshape(128, 64).t() + p[_GCN + 'get_s.s2.cnn.bias']
    s = torch.softmax(th @ ph.transpose(-1, -2), dim=-1)               # (B, T, V, V), no 1/sqrt(d)
    y = (s @ e) @ p[_GCN + 'weight']                                   # (B, T, V, 128)
    return y.permute(0, 3, 2, 1).contiguous()                          # (B, 128, V, T)


def gru_step(x_gates: Tensor, h: Tensor, w_hh: Tensor, b_hh: Tensor) -> Tensor:
    """One GRU update given input pre-activations ``x_gates = W_ih x + b_ih``; gate order r, z, n
    (torch.nn.GRU / GRUCell as used at vhoi/models.py:267,274,299 and :294-295,:319-320)."""
    hg = h @ w_hh.t() + b_hh
    xr, xz, xn = x_gates.chunk(3, dim=-1)
    hr, hz, hn = hg.chunk(3, dim=-1)
    r = torch.sigmoid(xr + hr)
    z = torch.sigmoid(xz + hz)
    n = torch.tanh(xn + r * hn)
    return (1.0 - z) * n + z * h


def bigru(p: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """Bidirectional single-layer GRU, batch_first, zero initial state, no packing.
    x: (R, T, D) -> (R, T, 2D).  Reference: nn.GRU called per entity at vhoi/models.py:997-1000."""
    R, T, D = x.shape
    outs = []
    for suffix, order in (('', range(T)), ('_reverse', range(T - 1, -1, -1))):
        w_ih, w_hh = p[f'{prefix}.weight_ih_l0{suffix}'], p[f'{prefix}.weight_hh_l0{suffix}']
        b_ih, b_hh = p[f'{prefix}.bias_ih_l0{suffix}'], p[f'{prefix}.bias_hh_l0{suffix}']
        gi = x @ w_ih.t() + b_ih
        h = x.new_zeros(R, D)
        hs = [None] * T
        for t in order:
            h = gru_step(gi[:, t], h, w_hh, b_hh)
            hs[t] = h
        outs.append(torch.stack(hs, dim=1))
    return torch.cat(outs, dim=-1)


def frame_level_rnn(p, x: Tensor, rnn: str, mlp: str) -> Tuple[Tensor, Tensor]:
    """_process_frame_level_rnn, vhoi/models.py:983-1002.  x: (B,T,E,D) -> h_f (B,T,E,D), h_fr (B,T,E,2D)."""
    h_fr = torch.stack([bigru(p, rnn, x[:, :, e]) for e in range(x.size(2))], dim=2)
    return _relu_lin(p, mlp + '.0', h_fr), h_fr


def attend(query: Tensor, keys: Tensor, values: Tensor, mask: Tensor, mean_pool: bool = False, scaled: bool = True,
           dist: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Scaled dot-product attention of compute_attention_weights (vhoi/models.py:1721-1754, style 'v3')
    followed by the weighted sum at e.g. :1047-1048.
    query (B,F), keys (B,S,F), values (B,S,D), mask (B,S) in {0,1}.  Fully masked rows give zeros
    (softmax of all -inf is NaN, replaced by 0 at :1753)."""
    if mean_pool:       # 'mp': sum of the (masked) messages over clamp(#valid senders, min=1), e.g. models.py:1033-1036
        w = mask / torch.clamp(mask.sum(dim=1, keepdim=True), min=1.0)
        return (w[..., None] * values).sum(1), w
    if dist is not None:        # compute_distance_based_attention_weights, models.py:1757-1775: softmax of 1 / (d + 1e-7) over the
        logits = 1 / (dist + 1e-7)                                      # real senders at a non-zero distance
        mask = mask * dist.bool().to(mask.dtype)
    else:
        logits = (query[:, None, :] * keys).sum(-1)
        if scaled:      # 'v3'; 'v2' is the plain dot product (models.py:1742-1745)
            logits = logits / math.sqrt(keys.size(-1))
    logits = torch.where(mask.bool(), logits, torch.full_like(logits, float('-inf')))
    w = torch.softmax(logits, dim=1)
    w = torch.where(torch.isnan(w), torch.zeros_like(w), w)
    return (w[..., None] * values).sum(1), w


def _others(x: Tensor, i: int) -> Tensor:
    return torch.cat([x[:, :i], x[:, i + 1:]], dim=1)


def gumbel_sigmoid(prob: Tensor, g: Optional[Tensor]) -> Tensor:
    """sample_from_gumbel_sigmoid, pyrutils/torch/distributions.py:15-18 (temperature 1).
    ``g`` is the (B, 2) Gumbel(0,1) draw; if None it is drawn exactly like the reference does."""
    pp = torch.cat([prob, 1.0 - prob], dim=-1)
    if g is None:
        g = torch.distributions.gumbel.Gumbel(0.0, 1.0).sample(pp.size())
    y = torch.log(pp + 1e-20) + g.to(pp)
    return torch.softmax(y, dim=-1)[:, :1]


def sample_gate(prob: Tensor, g: Optional[Tensor], thr: float, straight_through: bool) -> Tuple[Tensor, Tensor]:
    """discrete_estimator, vhoi/models.py:1620-1627: (hard, soft) decision of one entity at one step."""
    if straight_through:            # StraightThroughEstimator, distributions.py:39-53: exact 0/1 forward, identity backward
        return (prob > thr).to(prob.dtype) + (prob - prob.detach()), prob
    y = gumbel_sigmoid(prob, g)
    return hard_gate(y, thr), y


def hard_gate(y: Tensor, thr: float) -> Tensor:
    """straight_through_gumbel_sigmoid, pyrutils/torch/distributions.py:33-36: value of (z - y) + y."""
    z = (y > thr).to(y.dtype)
    return (z - y).detach() + y           # forward value z (up to rounding), d/dy = 1


def filter_soft(y: Tensor, thr: float) -> Tensor:
    """filter_soft_decisions, vhoi/models.py:1637-1664, on a (B, T) tensor of soft gates of one entity."""
    B, T = y.shape
    zero = y.new_zeros(B, 1)
    prev = torch.cat([zero, y[:, :-1]], dim=1)
    nxt = torch.cat([y[:, 1:], zero], dim=1)
    keep = (y > prev) & (y > nxt) & (y >= thr)
    u = ((y >= thr).to(y.dtype) - y).detach() + y      # straight-through, models.py:1660-1661
    return torch.where(keep, u, torch.clamp(u, max=0.0))


def reorder(hx: Tensor, u: Tensor) -> Tensor:
    """reorder_hidden_states, vhoi/models.py:1567-1586.  hx (B,T,F), u (B,T): every frame before a
    segment end takes the state of that end frame; frames after the last end keep their own."""
    B, T, _ = hx.shape
    out = hx.clone()
    for b in range(B):
        nxt = -1
        for t in range(T - 1, -1, -1):
            if u[b, t] != 0:
                nxt = t
            if nxt >= 0:
                out[b, t] = hx[b, nxt]
    return out


# ----------------------------------------------------------------------------------------------
# whole forward
# ----------------------------------------------------------------------------------------------
def time_embedding(p: Dict[str, Tensor], cfg: OracleConfig, steps_per_example: Tensor, T: int) -> Tensor:
    """(T, B, D) time-position features: _assemble_time_tensor (vhoi/models.py:936-952) followed by time_position_mlp
    (:259-260, embedding) or make_periodic_embedding (:1777-1794; no division by the number of steps, :654)."""
    D = cfg.hidden_size
    steps = steps_per_example.to(p['human_embedding_mlp.0.weight'].dtype)
    x = torch.arange(1, T + 1, dtype=steps.dtype).unsqueeze(-1).repeat_interleave(steps.size(0), dim=1)     # (T, B)
    if cfg.positional_encoding == 'e':
        return _relu_lin(p, 'time_position_mlp.0', (x / steps).unsqueeze(-1))
    w = torch.tensor([1e4], dtype=steps.dtype) ** torch.linspace(0, 1, D // 2, dtype=steps.dtype)
    x = x.unsqueeze(-1)
    return torch.cat([torch.sin(x / w), torch.cos(x / w)], dim=-1)


def _gate_prob(p: Dict[str, Tensor], cfg: OracleConfig, mlp: str, x: Tensor) -> Tensor:
    """update_*_segment_mlp (build_mlp, vhoi/models.py:532-547): gate_layers - 1 hidden Linear + ReLU layers, then Linear(., 1) + sigmoid;
    nn.Sequential indices 0, 2, 4, ..."""
    for i in range(cfg.gate_layers - 1):
        x = _relu_lin(p, f'{mlp}.{2 * i}', x)
    return torch.sigmoid(_lin(p, f'{mlp}.{2 * (cfg.gate_layers - 1)}', x))


def _embed_positions(p: Dict[str, Tensor], cfg: OracleConfig, x: Tensor, mlp: str) -> Tensor:
    """(…, 1) position values -> (…, D): Linear(1, D) + ReLU ('e') or make_periodic_embedding ('p', models.py:1777-1794)."""
    if cfg.positional_encoding == 'e':
        return _relu_lin(p, mlp, x)
    w = torch.tensor([1e4], dtype=x.dtype) ** torch.linspace(0, 1, cfg.hidden_size // 2, dtype=x.dtype)
    return torch.cat([torch.sin(x / w), torch.cos(x / w)], dim=-1)


def segment_lengths(u: Tensor, steps_per_example: Tensor, periodic: bool) -> Tensor:
    """_assemble_segment_length_tensor, vhoi/models.py:954-979, for the hard gates u (B, T, E): at a frame that closes a segment
    (u != 0) the time since the previous boundary, else u * time (= 0 in value; the hard gates keep their gradient)."""
    B, T, E = u.shape
    steps = steps_per_example.to(u.dtype)
    out = []
    acc = u.new_zeros(B, E)
    for t in range(T):
        x_t = u.new_full((B, 1), float(t + 1))
        if not periodic:
            x_t = x_t / steps[:, None]
        rel = u[:, t] * x_t
        rel = torch.where(rel.bool(), rel - acc, rel)
        acc = acc + rel
        out.append(rel)
    return torch.stack(out, dim=1)                                      # (B, T, E)


def forward(p: Dict[str, Tensor], cfg: OracleConfig, x_human: Tensor, x_objects: Tensor, objects_mask: Tensor,
            human_segmentation: Optional[Tensor] = None, objects_segmentation: Optional[Tensor] = None,
            noise: Optional[Tensor] = None, training: bool = False, inspect_model: bool = False,
            taps: Optional[dict] = None, gates_only: bool = False, steps_per_example: Optional[Tensor] = None,
            distances: Optional[Tuple[Optional[Tensor], Optional[Tensor], Optional[Tensor]]] = None):
    """TGGCN.forward, vhoi/models.py:584-933, for the shipped configuration family
    (message_type v2, granularity v1, attention aggregation style v3, update strategy 'ind',
    gumbel-sigmoid gates, message_segment on, geometry->objects on, geometry->human off).

    ``noise``: (n_calls, B, 2) Gumbel(0,1) draws consumed in the reference's call order
    (t-major; humans then objects; only entities whose segmentation is not given), or None to draw
    from the global CPU generator per call like pyrutils/torch/distributions.py:16.
    ``taps``: optional dict that receives named intermediates (kernel-level parity tests).
    ``gates_only``: stop after the frame-level part and return (y_hs, y_hss, y_os, y_oss) — used by the seed search of the
    full-size parity cases (tools/find_safe_seeds.py), which only needs the gate margins."""
    D, V, thr = cfg.hidden_size, cfg.gcn_node, cfg.update_segment_threshold
    hh_on = cfg.message_humans_to_human
    B, T, H, _ = x_human.shape
    O = x_objects.size(2)
    tap = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)
    noise_it = iter(noise) if noise is not None else None
    d_hh, d_ho, d_oo = distances if distances is not None else (None, None, None)      # (B,T,H,H), (B,T,H,O), (B,T,O,O) or None
    dist_hh = lambda t, h: None if d_hh is None else _others(d_hh[:, t, h], h)          # models.py:1044-1045
    dist_oh = lambda t, h: None if d_ho is None else d_ho[:, t, h]                      # receiver human h, senders objects (:683)
    dist_ho = lambda t, k: None if d_ho is None else d_ho[:, t, :, k]                   # receiver object k, senders humans (:715)
    dist_oo = lambda t, k: None if d_oo is None else _others(d_oo[:, t, k], k)          # :1327-1328

    # -- split, geometry GCN, scrambled view (models.py:631-645) -------------------------------
    vis = x_human[..., :2048]
    geo = x_human[:, :, 0, 2048:2048 + 4 * V]                          # human 0 only
    y = geo_gcn(p, geo.reshape(B, T, V, 4).permute(0, 3, 2, 1).contiguous(), training, taps)
    tap('gcn_out', y)
    x_g = y.reshape(B, T, 1, 128 * V)                                  # reinterpretation, NOT a permute
    # -- embeddings (models.py:646) -----------------------------------------------------------
    x_h = _relu_lin(p, 'human_embedding_mlp.0', vis)
    x_o = _relu_lin(p, 'object_embedding_mlp.0', x_objects)
    x_g = _relu_lin(p, 'geometry_embedding_mlp.2', _relu_lin(p, 'geometry_embedding_mlp.0', x_g))
    tap('x_h', x_h), tap('x_o', x_o), tap('x_g', x_g)
    # -- frame-level BiGRUs (models.py:649-651) -----------------------------------------------
    h_h, hfr_h = frame_level_rnn(p, x_h, 'human_bd_rnn', 'human_bd_embedding_mlp')
    h_o, hfr_o = frame_level_rnn(p, x_o, 'object_bd_rnn', 'object_bd_embedding_mlp')
    h_g, hfr_g = frame_level_rnn(p, x_g, 'geometry_bd_rnn', 'geometry_bd_embedding_mlp')
    tap('hfr_h', hfr_h), tap('hfr_o', hfr_o), tap('hfr_g', hfr_g)
    tap('h_h', h_h), tap('h_o', h_o), tap('h_g', h_g)

    # -- frame-level messages and gates, one timestep at a time (models.py:664-749) ------------
    ones_h = x_h.new_ones(B, max(H - 1, 0))
    ones_H = x_h.new_ones(B, H)
    xx_h = [[None] * T for _ in range(H)]
    xx_o = [[None] * T for _ in range(O)]
    hard_h = [[None] * T for _ in range(H)]
    soft_h = [[None] * T for _ in range(H)]
    hard_o = [[None] * T for _ in range(O)]
    soft_o = [[None] * T for _ in range(O)]
    att_oh = [[None] * T for _ in range(H)]
    tt = time_embedding(p, cfg, steps_per_example, T) if cfg.time_position else None     # (T,B,D)
    for t in range(T):
        s_h = torch.cat([x_h[:, t], h_h[:, t]], dim=-1)                # (B,H,2D) sender/receiver features
        s_o = torch.cat([x_o[:, t], h_o[:, t]], dim=-1)
        s_g = torch.cat([x_g[:, t], h_g[:, t]], dim=-1)
        for h in range(H):
            parts = [h_h[:, t, h]]
            gate_in = [x_h[:, t, h], h_h[:, t, h]]
            if hh_on:                                                   # :1004-1049
                snd = _others(s_h, h)
                m_hh, _ = attend(s_h[:, h], snd, _relu_lin(p, 'humans_to_human_message_mlp.0', snd), ones_h, cfg.mean_pool, cfg.att_scaled, dist_hh(t, h))
                parts.append(m_hh), gate_in.append(m_hh)
            val = _relu_lin(p, 'objects_to_human_message_mlp.0', s_o) * objects_mask[..., None]   # :1191-1237
            m_oh, w_oh = attend(s_h[:, h], s_o, val, objects_mask, cfg.mean_pool, cfg.att_scaled, dist_oh(t, h))
            att_oh[h][t] = w_oh
            parts.append(m_oh), gate_in.append(m_oh)
            if cfg.geo_to_human:                                        # :690-695, :1432-1475: single sender, weight 1, no mask
                m_gh = _relu_lin(p, 'geometry_to_human_message_mlp.0', s_g[:, 0])
                parts.append(m_gh), gate_in.append(m_gh)
            if human_segmentation is not None:                          # :697-698
                hard_h[h][t] = soft_h[h][t] = human_segmentation[:, t:t + 1, h]
            else:                                                       # :1477-1498, :700-702
                if cfg.time_position == 'u':
                    gate_in.append(tt[t])
                prob = _gate_prob(p, cfg, 'update_human_segment_mlp', torch.cat(gate_in, dim=-1))
                z, ysoft = sample_gate(prob, next(noise_it) if (noise_it is not None and not cfg.straight_through) else None, thr, cfg.straight_through)
                if t == T - 1:
                    z = torch.ones_like(z)
                hard_h[h][t], soft_h[h][t] = z, ysoft
            if cfg.time_position == 's':
                parts.append(tt[t])                                     # :755-762 (appended after the frame loop there)
            xx_h[h][t] = torch.cat(parts, dim=-1)                       # :705  [h, m_hh, m_oh]
        for k in range(O):
            mk = objects_mask[:, k:k + 1]
            m_ho, _ = attend(s_o[:, k], s_h, _relu_lin(p, 'human_to_object_message_mlp.0', s_h), ones_H, cfg.mean_pool, cfg.att_scaled, dist_ho(t, k))
            m_ho = m_ho * mk                                            # :1099-1143, :720
            m_go = _relu_lin(p, 'geometry_to_object_message_mlp.0', s_g[:, 0]) * mk   # :1384-1428, :729
            snd, snd_mask = _others(s_o, k), _others(objects_mask, k)   # :1286-1332
            val = _relu_lin(p, 'objects_to_object_message_mlp.0', snd) * snd_mask[..., None]
            m_oo, _ = attend(s_o[:, k], snd, val, snd_mask, cfg.mean_pool, cfg.att_scaled, dist_oo(t, k))
            if taps is not None:                                        # per-(t, receiver) tensors for gradient debugging
                taps.setdefault('val_oo', {})[(t, k)] = val
                taps.setdefault('m_oo', {})[(t, k)] = m_oo
            if objects_segmentation is not None:                        # :738-739
                hard_o[k][t] = soft_o[k][t] = objects_segmentation[:, t:t + 1, k]
            elif cfg.update_strategy == 'sah' and H == 1:               # :741-742, :1523-1525: the human's decision, no object MLP
                hard_o[k][t], soft_o[k][t] = hard_h[0][t], soft_h[0][t]
            else:                                                       # :1500-1533 ('ind' / 'coh'); input order :1527
                gate_in = torch.cat([x_o[:, t, k], h_o[:, t, k], m_ho, m_oo, m_go] + ([tt[t]] if cfg.time_position == 'u' else []), dim=-1)
                prob = _gate_prob(p, cfg, 'update_object_segment_mlp', gate_in)
                z, ysoft = sample_gate(prob, next(noise_it) if (noise_it is not None and not cfg.straight_through) else None, thr, cfg.straight_through)
                if cfg.update_strategy == 'coh' and H == 1:             # :1531-1532: object updates only where the human does
                    z = z * hard_h[0][t]
                if t == T - 1:
                    z = torch.ones_like(z)
                hard_o[k][t], soft_o[k][t] = z, ysoft
            xx_o[k][t] = torch.cat([h_o[:, t, k], m_ho, m_go, m_oo] + ([tt[t]] if cfg.time_position == 's' else []), dim=-1)    # :748
    y_hss = torch.stack([torch.cat(s, dim=-1) for s in soft_h], dim=-1)          # (B,T,H)
    y_oss = torch.stack([torch.cat(s, dim=-1) for s in soft_o], dim=-1)          # (B,T,O)
    y_hs = torch.stack([torch.cat(s, dim=-1) for s in hard_h], dim=-1)
    y_os = torch.stack([torch.cat(s, dim=-1) for s in hard_o], dim=-1)
    # -- optional local-maximum filter replaces the hard gates (models.py:751-753) -------------
    if cfg.filter_discrete_updates:
        y_hs = torch.stack([filter_soft(y_hss[..., h], thr) for h in range(H)], dim=-1)
        y_os = torch.stack([filter_soft(y_oss[..., k], thr) for k in range(O)], dim=-1)
    if gates_only:
        return y_hs, y_hss, y_os, y_oss
    if cfg.segment_length:                                                      # models.py:763-779
        for hard, xx in ((y_hs, xx_h), (y_os, xx_o)):
            emb = _embed_positions(p, cfg, segment_lengths(hard, steps_per_example, cfg.positional_encoding == 'p').unsqueeze(-1),
                                   'segment_length_mlp.0')                      # (B,T,E,D)
            for e in range(len(xx)):
                for t in range(T):
                    xx[e][t] = torch.cat([xx[e][t], emb[:, t, e]], dim=-1)
    tap('xx_h', torch.stack([torch.stack(r, dim=1) for r in xx_h], dim=2))       # (B,T,H,3D)
    tap('xx_o', torch.stack([torch.stack(r, dim=1) for r in xx_o], dim=2))       # (B,T,O,4D)

    # -- segment-level recurrent graph, both directions in lock-step (models.py:785-880) -------
    def seg_dir(order, hcell: str, ocell: str):
        SH = [x_h.new_zeros(B, D) for _ in range(H)]
        SO = [x_h.new_zeros(B, D) for _ in range(O)]
        out_h = [[None] * T for _ in range(H)]
        out_o = [[None] * T for _ in range(O)]
        att = [[None] * T for _ in range(H)]
        for t in order:
            sh, so = torch.stack(SH, dim=1), torch.stack(SO, dim=1)     # previous-step states only
            newH, newO = [], []
            for h in range(H):
                x = [xx_h[h][t]]
                if hh_on:                                               # :1051-1097
                    snd = _others(sh, h)
                    mg, _ = attend(SH[h], snd, _relu_lin(p, 'humans_to_human_segment_message_mlp.0', snd), ones_h, cfg.mean_pool, cfg.att_scaled, dist_hh(t, h))
                    x.append(mg)
                val = _relu_lin(p, 'objects_to_human_segment_message_mlp.0', so) * objects_mask[..., None]
                mg, w = attend(SH[h], so, val, objects_mask, cfg.mean_pool, cfg.att_scaled, dist_oh(t, h))            # :1239-1284
                att[h][t] = w
                x.append(mg)
                newH.append(_seg_cell(p, hcell, torch.cat(x, dim=-1), y_hs[:, t, h:h + 1], SH[h]))
            for k in range(O):
                mg_ho, _ = attend(SO[k], sh, _relu_lin(p, 'human_to_object_segment_message_mlp.0', sh), ones_H, cfg.mean_pool, cfg.att_scaled, dist_ho(t, k))
                snd, snd_mask = _others(so, k), _others(objects_mask, k)        # :1334-1381
                val = _relu_lin(p, 'objects_to_object_segment_message_mlp.0', snd) * snd_mask[..., None]
                mg_oo, _ = attend(SO[k], snd, val, snd_mask, cfg.mean_pool, cfg.att_scaled, dist_oo(t, k))
                x = torch.cat([xx_o[k][t], mg_ho, mg_oo], dim=-1)
                newO.append(_seg_cell(p, ocell, x, y_os[:, t, k:k + 1], SO[k]))
            SH, SO = newH, newO                                         # commit, :875-880
            for h in range(H):
                out_h[h][t] = SH[h]
            for k in range(O):
                out_o[k][t] = SO[k]
        oh = torch.stack([torch.stack(r, dim=1) for r in out_h], dim=2)         # (B,T,H,D)
        oo = torch.stack([torch.stack(r, dim=1) for r in out_o], dim=2)
        aw = torch.stack([torch.stack(r, dim=1) for r in att], dim=1)           # (B,H,T,O)
        return oh, oo, aw

    fh, fo, att_f = seg_dir(range(T), 'human_segment_rnn_fcell', 'object_segment_rnn_fcell')
    bh, bo, att_b = seg_dir(range(T - 1, -1, -1), 'human_segment_rnn_bcell', 'object_segment_rnn_bcell')
    hx_h = torch.cat([fh, bh], dim=-1)                                          # (B,T,H,2D)
    hx_o = torch.cat([fo, bo], dim=-1)
    tap('hx_h', hx_h), tap('hx_o', hx_o)
    # -- reorder (models.py:886-899) -------------------------------------------------------------
    hx_h = torch.stack([reorder(hx_h[:, :, h], y_hs[:, :, h]) for h in range(H)], dim=2)
    hx_o = torch.stack([reorder(hx_o[:, :, k], y_os[:, :, k]) for k in range(O)], dim=2)
    tap('hx_h_reordered', hx_h), tap('hx_o_reordered', hx_o)

    if cfg.cat_level_states:                                                    # models.py:901-903
        hx_h = torch.cat([hx_h, hfr_h], dim=-1)
        hx_o = torch.cat([hx_o, hfr_o], dim=-1)
    # -- heads (models.py:905-926) ---------------------------------------------------------------
    def head(name, x):
        return torch.log_softmax(_lin(p, name + '.0', x), dim=-1).permute(0, 3, 1, 2).contiguous()

    out_h = [head('human_frame_recognition_mlp', hfr_h), head('human_frame_prediction_mlp', hfr_h),
             head('human_recognition_mlp', hx_h), head('human_prediction_mlp', hx_h)]
    if cfg.num_classes[1] is None:
        output = [y_hs, y_hss] + out_h
    else:
        out_o = [head('object_frame_recognition_mlp', hfr_o), head('object_frame_prediction_mlp', hfr_o),
                 head('object_recognition_mlp', hx_o), head('object_prediction_mlp', hx_o)]
        output = [y_hs, y_os, y_hss, y_oss] + out_h[:2] + out_o[:2] + out_h[2:] + out_o[2:]
    tap('y_os', y_os), tap('y_oss', y_oss)
    if inspect_model:
        ax_hf = torch.stack([torch.stack(r, dim=1) for r in att_oh], dim=1)     # (B,H,T,O)  :928
        return output, [ax_hf, att_f, att_b]
    return output


def _seg_cell(p, cell: str, x: Tensor, u: Tensor, h: Tensor) -> Tensor:
    """_bidirectional_step, vhoi/models.py:1535-1564: u * GRUCell(x, h) + (1 - u) * h."""
    gi = x @ p[cell + '.weight_ih'].t() + p[cell + '.bias_ih']
    new = gru_step(gi, h, p[cell + '.weight_hh'], p[cell + '.bias_hh'])
    return u * new + (1.0 - u) * h


def num_noise_draws(T: int, H: int, O: int, human_given: bool, objects_given: bool, update_strategy: str = 'ind',
                    straight_through: bool = False) -> int:
    """Number of (B,2) Gumbel draws one forward consumes (vhoi/models.py:697-702, :738-745); under 'sah' with one human the
    objects copy the human's decision and draw nothing (:1523-1525)."""
    if straight_through:            # discrete_optimization_strategy 'st' samples nothing
        return 0
    objects_sampled = not objects_given and not (_UPD[update_strategy] == 'sah' and H == 1)
    return T * ((0 if human_given else H) + (O if objects_sampled else 0))


def draw_noise(n_calls: int, B: int, generator: Optional[torch.Generator] = None) -> Tensor:
    """Pre-draw the Gumbel(0,1) noise of a forward in one call.  Equal, draw for draw, to what
    torch.distributions.gumbel.Gumbel(0,1).sample((B,2)) yields call by call:
    Gumbel.sample = -log(-log(U)), U = rand()*(1-eps-tiny)+tiny."""
    fi = torch.finfo(torch.float32)
    u = torch.rand(n_calls, B, 2, generator=generator)
    u = u * ((1.0 - fi.eps) - fi.tiny) + fi.tiny
    return -torch.log(-torch.log(u))


# ----------------------------------------------------------------------------------------------
# losses (pyrutils/torch/losses.py:7-51, vhoi/losses.py:8-61)
# ----------------------------------------------------------------------------------------------
def budget_loss(inp: Tensor, tgt: Tensor) -> Tensor:
    """budget_loss, pyrutils/torch/losses.py:24-36."""
    mask = (tgt != -1).to(inp.dtype)
    n = float(mask.sum())
    if n == 0:
        return inp.new_zeros(())
    return (inp * mask).mean() * (inp.numel() / n)


def bce_loss(inp: Tensor, tgt: Tensor) -> Tensor:
    """binary_cross_entropy_loss, pyrutils/torch/losses.py:7-21 (positive_class_weight == 1).
    F.binary_cross_entropy clamps each log term at -100."""
    mask = (tgt != -1).to(inp.dtype)
    n = float(mask.sum())
    if n == 0:
        return inp.new_zeros(())
    o, t = inp * mask, tgt * mask
    # = mean(-[t*max(log o, -100) + (1-t)*max(log(1-o), -100)]); the library op also has the finite backward at o == 0
    return torch.nn.functional.binary_cross_entropy(o, t, reduction='mean') * (inp.numel() / n)


def nll_loss(logp: Tensor, tgt: Tensor) -> Tensor:
    """F.nll_loss(ignore_index=-1, reduction='mean') on (B,C,T,E) log-probs and (B,T,E) int64 targets
    (pyrutils/torch/losses.py:46-47)."""
    valid = tgt != -1
    picked = torch.gather(logp, 1, tgt.clamp(min=0).unsqueeze(1)).squeeze(1)
    return -(picked * valid.to(logp.dtype)).sum() / valid.sum().to(logp.dtype)


def loss_weights(dataset: str, stage: int) -> List[float]:
    """select_loss weights for the shipped yaml files (vhoi/losses.py:8-61 with
    conf/models/2G-GCN_stage{1,2}.yaml:37-54): stage 2 switches the segmentation BCE on."""
    s = 1.0 if stage == 2 else 0.0
    if dataset == 'cad120':
        return [0.0, 0.0, s, s, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0]
    return [0.0, s, 0.0, 0.0, 1.0, 1.0]


def multi_task_loss(outputs: List[Tensor], targets: List[Tensor], dataset: str, stage: int) -> List[Tensor]:
    """multi_task_loss, pyrutils/torch/losses.py:39-51, with the function tuple of vhoi/losses.py:41-60."""
    if dataset == 'cad120':
        fns = [budget_loss, budget_loss, bce_loss, bce_loss] + [nll_loss] * 8
    else:
        fns = [budget_loss, bce_loss] + [nll_loss] * 4
    return [w * fn(o, t) for o, t, fn, w in zip(outputs, targets, fns, loss_weights(dataset, stage))]


# ----------------------------------------------------------------------------------------------
# F1@k (pyrutils/metrics.py:7-81, pyrutils/utils.py:56-60, pyrutils/itertools.py:15-18)
# ----------------------------------------------------------------------------------------------
def _runs(seq):
    labels, lengths = zip(*[(k, len(list(v))) for k, v in groupby(seq)])
    starts = [0] + list(accumulate(lengths))
    return np.array(labels), np.array(list(zip(starts[:-1], starts[1:])))


def f1_at_k_single(y_true, y_pred, num_classes: int, overlap: float) -> float:
    """f1_at_k_single_example, pyrutils/metrics.py:7-61."""
    t_ids, t_iv = _runs(y_true)
    o_ids, o_iv = _runs(y_pred)
    tp = fp = 0.0
    used = np.zeros(len(t_ids), dtype=bool)
    for (a, b), oid in zip(o_iv, o_ids):
        inter = np.minimum(b, t_iv[:, 1]) - np.maximum(a, t_iv[:, 0])
        union = np.maximum(b, t_iv[:, 1]) - np.minimum(a, t_iv[:, 0])
        iou = (inter / union) * (oid == t_ids)
        j = int(np.argmax(iou))
        if oid >= num_classes:
            continue
        if iou[j] >= overlap and not used[j]:
            tp += 1
            used[j] = True
        else:
            fp += 1
    fn = len(used) - float(used.sum())
    prec = tp / (tp + fp) if (tp + fp) else 0.0
    rec = tp / (tp + fn) if (tp + fn) else 0.0
    return 2 * prec * rec / (prec + rec) if (prec + rec) else 0.0


def f1_at_k(y_true, y_pred, num_classes: int, overlap: float, ignore_value=-1.0) -> float:
    """f1_at_k, pyrutils/metrics.py:64-81.  y_* are (rows, T) label arrays."""
    total, n = 0.0, 0
    for yt, yp in zip(np.asarray(y_true), np.asarray(y_pred)):
        keep = yt != ignore_value
        yt, yp = yt[keep], yp[keep]
        if yt.size == 0:
            continue
        total += f1_at_k_single(yt, yp, num_classes, overlap)
        n += 1
    return total / n


def labels_for_f1(arr: np.ndarray) -> np.ndarray:
    """The reshape convention of predict.py:236-240: (B,T,E) -> swapaxes(1,2) -> (B*E, T)."""
    if arr.ndim == 3:
        arr = np.swapaxes(arr, 1, 2)
    return arr.reshape(-1, arr.shape[-1])
