"""Generate the golden vectors under tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference, which does not
exist on the GPU box); its outputs are committed so nothing at test/bench time reads the reference.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

For every case the reference ``TGGCN`` (vhoi/models.py:178) is built with the yaml constructor
arguments, its ``state_dict`` is overwritten by ``synth.deterministic_fill`` (values depend only on
seed/key/shape, so any box can regenerate the same weights), the Gumbel sampler
(pyrutils/torch/distributions.py:4) is replaced by one that consumes a pre-drawn noise tensor in call
order, and the outputs of ``model(**kwargs)`` — called through the reference's own ``gcn_forward``
(vhoi/data_loading.py:1233) — are stored together with the reference's losses
(vhoi/losses.py:8 → pyrutils/torch/losses.py:39) and F1@k (pyrutils/metrics.py:64).
Seeds are advanced until every sampled gate is at least 1e-4 away from its threshold / its
neighbours, so that fp32 re-association in another implementation cannot flip a discrete decision.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('2g-gcn_b200.synth')

sys.modules.setdefault('zarr', types.ModuleType('zarr'))   # data_loading imports zarr at top (:14)
sys.path.insert(0, '/root/reference')
from vhoi.models import select_model                                  # noqa: E402
from vhoi.data_loading import select_model_data_feeder                # noqa: E402
from vhoi.losses import select_loss                                   # noqa: E402
from pyrutils.metrics import f1_at_k                                  # noqa: E402
import pyrutils.torch.distributions as ref_dist                      # noqa: E402

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import tggcn_oracle as orc                                            # noqa: E402


def dists64(dists):
    return None if dists is None else tuple(None if d is None else d.double() for d in dists)


class Cfg(dict):
    def get(self, k, default_value=None):
        return dict.get(self, k, default_value)


CASES = [
    # name, shape, D, B, T, stage, train_mode, gain, inspect
    ('mphoi_s1_eval', 'mphoi', 32, 2, 12, 1, False, 2.0, False),
    ('mphoi_s2_eval', 'mphoi', 32, 3, 14, 2, False, 2.0, True),
    ('mphoi_s2_train_bn', 'mphoi', 32, 2, 11, 2, True, 2.0, False),
    ('mphoi_s2_d64', 'mphoi', 64, 4, 40, 2, False, 1.5, False),
    ('cad120_s1_eval', 'cad120', 32, 2, 10, 1, False, 2.0, False),
    ('cad120_s2_eval', 'cad120', 32, 2, 13, 2, False, 2.0, False),
    ('bimanual_s2_eval', 'bimanual', 16, 2, 9, 2, False, 2.0, False),
    # model variants beyond the shipped yaml values: trailing dict = constructor overrides
    ('mphoi_s2_cat', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'cat_level_states': 1}),
    ('cad120_s2_cat', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'cat_level_states': 1}),
    ('mphoi_s2_share', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'share_level_mlps': 1}),
    ('mphoi_s2_mp', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'message_aggregation': 'mp'}),
    ('cad120_s2_mp', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'message_aggregation': 'mp'}),
    ('mphoi_s2_dot', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'attention_style': 'v2'}),
    ('cad120_s2_dot', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'attention_style': 'v2'}),
    # object_segment_update_strategy (models.py:1523-1532; acts with exactly one human = CAD-120).  Under the local-maximum
    # filter 'coh' equals 'ind' (the filter recomputes the hard gates from the soft ones, :751-753), so it is pinned without it.
    ('cad120_s2_sah', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'object_segment_update_strategy': 'sah'}),
    ('cad120_nf_sah', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'object_segment_update_strategy': 'sah', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    ('cad120_nf_coh', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'object_segment_update_strategy': 'coh', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    # add_time_position (models.py:259-260, :656-662, :755-762): strategy 's' (segment-level input) / 'u' (gate input), encoding e / p
    ('mphoi_s2_time_se', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'add_time_position': 1, 'time_position_strategy': 's', 'positional_encoding_style': 'e'}),
    ('cad120_s2_time_sp', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'add_time_position': 1, 'time_position_strategy': 's', 'positional_encoding_style': 'p'}),
    ('cad120_s2_time_ue', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'add_time_position': 1, 'time_position_strategy': 'u', 'positional_encoding_style': 'e'}),
    ('mphoi_s2_time_up', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'add_time_position': 1, 'time_position_strategy': 'u', 'positional_encoding_style': 'p'}),
    # discrete_optimization_strategy 'st' (models.py:1620-1622): no Gumbel noise, soft gate = sigmoid probability
    ('mphoi_s2_st', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'discrete_optimization_strategy': 'st'}),
    ('cad120_nf_st', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'discrete_optimization_strategy': 'st', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    # message_geometry_to_human (models.py:690-695, :1432-1475), also together with a time block in the gate inputs
    ('mphoi_s2_gh', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'message_geometry_to_human': True}),
    ('cad120_s2_gh_time_u', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'message_geometry_to_human': True, 'add_time_position': 1, 'time_position_strategy': 'u'}),
    # add_segment_length (models.py:763-779, :954-979): embedding / periodic, with and without the filter, and after a time block
    ('mphoi_s2_len_e', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'add_segment_length': 1}),
    ('cad120_nf_len_p', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'add_segment_length': 1, 'positional_encoding_style': 'p', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    ('cad120_s2_time_len', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'add_segment_length': 1, 'add_time_position': 1}),
    # misc.make_attention_distance_based (data_loading.py:1264-1276; compute_distance_based_attention_weights, models.py:1757-1775)
    ('mphoi_s2_dist', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'_distances': True}),
    ('cad120_s2_dist', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'_distances': True}),
    # discrete_networks_num_layers = 2 (models.py:532-547): hidden layer in the gate MLPs; alone and with every gate-input block + 'coh'
    ('mphoi_s2_gate2', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'discrete_networks_num_layers': 2}),
    ('cad120_nf_gate2_mix', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'discrete_networks_num_layers': 2, 'add_time_position': 1, 'time_position_strategy': 'u', 'message_geometry_to_human': True, 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5, 'object_segment_update_strategy': 'coh'}),
    ('mphoi_s2_gate3', 'mphoi', 32, 2, 12, 2, False, 2.0, False, {'discrete_networks_num_layers': 3}),
    ('cad120_s2_gate3_sah_u', 'cad120', 32, 2, 11, 2, False, 2.0, False, {'discrete_networks_num_layers': 3, 'add_time_position': 1, 'time_position_strategy': 'u', 'object_segment_update_strategy': 'sah'}),
    # the benchmarked configuration itself (BASELINE.json configs[1]: MPHOI, B=8, T=128, hidden 512, stage-2 settings)
    ('mphoi_s2_d512_full', 'mphoi', 512, 8, 128, 2, False, 1.0, False),
]

# Decision margin demanded of a case's seeds.  6144 gates at the full size leave no seed with 1e-4 everywhere; the CUDA path
# reproduces soft gates to ~1e-6, so 2e-5 still keeps every discrete decision off the knife edge.
CASE_MARGIN = {'mphoi_s2_d512_full': 2e-5, 'grad_mphoi_s2_d512': 2e-5}


def run_case(name, shape_name, D, B, T, stage, train_mode, gain, inspect, extra=None):
    shape = pkg.SHAPES[shape_name]
    extra = extra or {}
    kw = pkg.model_kwargs(shape, hidden_size=D, stage=stage, **extra)
    model = select_model('2G-GCN')(**kw)
    pkg.deterministic_fill(model.state_dict(), seed=7, gain=gain)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}   # snapshot (train mode mutates BN)
    model.train(train_mode)
    misc = dict(impose_segmentation_pattern=1 if stage == 1 else 0,
                segmentation_loss=dict(add=(stage == 2), sigma=4.0 if stage == 2 else 0.0, weight=1.0),
                make_attention_distance_based=bool(extra.get('_distances')))
    feed = select_model_data_feeder('2G-GCN', 'multiple', dataset_name=shape.dataset, inspect_model=inspect, **misc)
    criterion, _ = select_loss('2G-GCN', 'multiple', shape.dataset, cfg=Cfg(misc=misc))
    thr = kw['update_segment_threshold']
    for attempt in range(50):
        data_seed, noise_seed = 100 + attempt, 500 + attempt
        batch = pkg.make_batch(shape, B, T, seed=data_seed)
        dists = pkg.make_distances(shape, B, T, seed=data_seed + 5000) if extra.get('_distances') else None
        dist_data = [None, None, None] if dists is None else [None if shape.dataset == 'cad120' else dists[0], dists[1], dists[2]]
        n_calls = orc.num_noise_draws(T, shape.H, shape.O, stage == 1, stage == 1 and shape.dataset == 'cad120',
                                      kw['object_segment_update_strategy'],
                                      kw['discrete_optimization_strategy'] in ('st', 'straight-through'))
        noise = orc.draw_noise(max(n_calls, 1), B, torch.Generator().manual_seed(noise_seed))[:n_calls]
        if name in CASE_MARGIN and not train_mode:
            # big cases: pre-screen both human and object gates on the oracle's early exit before paying for the reference
            with torch.no_grad():
                _, s_h, _, s_o = orc.forward({k: v.double() for k, v in sd.items()},
                                             orc.config_from_kwargs(kw),
                                             batch['x_human'].double(), batch['x_objects'].double(), batch['objects_mask'].double(),
                                             None, None, noise.double(), gates_only=True, steps_per_example=batch['steps_per_example'], distances=dists64(dists))
            pre = 1.0
            for sft in (s_h, s_o):
                pre = min(pre, float((sft - thr).abs().min()), float((sft[:, 1:] - sft[:, :-1]).abs().min()))
            if pre <= CASE_MARGIN[name]:
                continue
        it = iter(noise)

        def injected(p, temperature=1.0):
            y = torch.log(torch.cat([p, 1.0 - p], -1) + 1e-20) + next(it).to(p)
            return torch.softmax(y / temperature, -1)[:, :1]

        ref_dist.sample_from_gumbel_sigmoid = injected
        model.load_state_dict(sd)
        captured = {}
        hk = model.geometry_embedding_gcn.register_forward_hook(lambda m, i, o: captured.__setitem__('gcn_out', o.detach().clone()))
        # data tuple layout of gcn_fetcher (data_loading.py:1282-1315)
        data = [batch['x_human'], batch['x_objects'], batch['objects_mask'], None] + dist_data + [batch['steps_per_example']]
        with torch.no_grad():
            res = feed(model, data)
        hk.remove()
        out, att = (res if inspect else (res, None))
        soft = [o for o in (out[1:2] if shape.num_classes[1] is None else out[2:4])]
        sampled = []
        if stage != 1:
            sampled = soft
        elif shape.dataset != 'cad120':
            sampled = []      # humans imposed; object gates are sampled but not returned for MPHOI
        margin = 1.0
        for s in sampled:
            margin = min(margin, float((s - thr).abs().min()))
            if kw['filter_discrete_updates']:
                margin = min(margin, float((s[:, 1:] - s[:, :-1]).abs().min()))
        if margin > CASE_MARGIN.get(name, 1e-4):
            break
    else:
        raise RuntimeError(f'no seed with a safe gate margin for {name}')
    # oracle agreement (also guards MPHOI object gates, which the reference does not return)
    p64 = {k: v.double() for k, v in sd.items()}
    ocfg = orc.config_from_kwargs(kw)
    hseg = torch.ones(B, T, shape.H) if stage == 1 else None
    oseg = torch.ones(B, T, shape.O) if (stage == 1 and shape.dataset == 'cad120') else None
    taps = {}
    o64 = orc.forward(p64, ocfg, batch['x_human'].double(), batch['x_objects'].double(),
                      batch['objects_mask'].double(), None if hseg is None else hseg.double(),
                      None if oseg is None else oseg.double(), noise.double(), training=train_mode, taps=taps,
                      steps_per_example=batch['steps_per_example'], distances=dists64(dists))
    o_margin = float((taps['y_oss'] - thr).abs().min()) if oseg is None else 1.0
    if o_margin <= CASE_MARGIN.get(name, 1e-4):
        raise RuntimeError(f'{name}: object gate margin too small ({o_margin}); change seeds')
    worst = max(float((a.double() - b).abs().max()) for a, b in zip(out, o64))
    print(f'{name:20s} seeds=({data_seed},{noise_seed}) gate margin={min(margin, o_margin):.2e} '
          f'|reference - oracle(fp64)|max={worst:.2e}')
    tg = pkg.make_targets(shape, batch['lengths'], T, seed=900 + attempt)
    targets = pkg.target_list(shape, tg)
    losses = criterion(out, targets, reduction='mean')
    # F1@k on the segment-level recognition output, predict.py:229-246 convention
    rec_idx = 4 if shape.num_classes[1] is None else 8
    pred = out[rec_idx].argmax(dim=1).numpy()                           # (B,T,H)
    tgt = tg['rec_h'].numpy()
    f1 = [f1_at_k(orc.labels_for_f1(tgt), orc.labels_for_f1(pred), shape.num_classes[0], overlap=k,
                  ignore_value=-1.0) for k in (0.10, 0.25, 0.50)]
    blob = {f'out{i}': o.numpy() for i, o in enumerate(out)}
    blob['losses'] = np.array([float(l) for l in losses], dtype=np.float64)
    blob['f1'] = np.array(f1, dtype=np.float64)
    if name not in CASE_MARGIN:          # 13.6 MB at the full size; the small cases pin the GCN output
        blob['gcn_out'] = captured['gcn_out'].numpy()
    blob['meta'] = np.array([data_seed, noise_seed, 900 + attempt, 7], dtype=np.int64)
    blob['gain'] = np.array([gain])
    blob['weights_checksum'] = np.array([pkg.state_checksum(sd)])
    blob['inputs_checksum'] = np.array([float(batch['x_human'].double().sum() + batch['x_objects'].double().sum())])
    blob['noise_checksum'] = np.array([float(noise.double().sum())])
    if train_mode:
        after = model.state_dict()
        for k in after:
            if '.bn.running' in k:
                blob['bn_after.' + k.rsplit('.', 1)[1]] = after[k].numpy()
    if att is not None:
        for i, a in enumerate(att):
            blob[f'att{i}'] = a.numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', name + '.npz'), **blob)


GRAD_CASES = [
    # name, shape, D, B, T, stage, gain  -- full training-mode forward + multi_task_loss + backward of the reference
    ('grad_mphoi_s1', 'mphoi', 32, 2, 9, 1, 2.0),
    ('grad_mphoi_s2', 'mphoi', 32, 3, 10, 2, 2.0),
    ('grad_cad120_s2', 'cad120', 32, 2, 8, 2, 2.0),
    ('grad_mphoi_s2_cat', 'mphoi', 32, 2, 9, 2, 2.0, {'cat_level_states': 1}),
    ('grad_cad120_s2_cat', 'cad120', 32, 2, 8, 2, 2.0, {'cat_level_states': 1}),
    ('grad_mphoi_s2_share', 'mphoi', 32, 2, 9, 2, 2.0, {'share_level_mlps': 1}),
    ('grad_mphoi_s2_mp', 'mphoi', 32, 2, 9, 2, 2.0, {'message_aggregation': 'mp'}),
    ('grad_cad120_s2_mp', 'cad120', 32, 2, 8, 2, 2.0, {'message_aggregation': 'mp'}),
    ('grad_mphoi_s2_dot', 'mphoi', 32, 2, 9, 2, 2.0, {'attention_style': 'v2'}),
    ('grad_cad120_s2_dot', 'cad120', 32, 2, 8, 2, 2.0, {'attention_style': 'v2'}),
    ('grad_cad120_s2_sah', 'cad120', 32, 2, 8, 2, 2.0, {'object_segment_update_strategy': 'sah'}),
    ('grad_cad120_nf_sah', 'cad120', 32, 2, 8, 2, 2.0, {'object_segment_update_strategy': 'sah', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    ('grad_cad120_nf_coh', 'cad120', 32, 2, 8, 2, 2.0, {'object_segment_update_strategy': 'coh', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    ('grad_mphoi_s2_time_se', 'mphoi', 32, 2, 9, 2, 2.0, {'add_time_position': 1, 'time_position_strategy': 's', 'positional_encoding_style': 'e'}),
    ('grad_cad120_s2_time_sp', 'cad120', 32, 2, 8, 2, 2.0, {'add_time_position': 1, 'time_position_strategy': 's', 'positional_encoding_style': 'p'}),
    ('grad_cad120_s2_time_ue', 'cad120', 32, 2, 8, 2, 2.0, {'add_time_position': 1, 'time_position_strategy': 'u', 'positional_encoding_style': 'e'}),
    ('grad_mphoi_s2_time_up', 'mphoi', 32, 2, 9, 2, 2.0, {'add_time_position': 1, 'time_position_strategy': 'u', 'positional_encoding_style': 'p'}),
    ('grad_mphoi_s2_gh', 'mphoi', 32, 2, 9, 2, 2.0, {'message_geometry_to_human': True}),
    ('grad_cad120_s2_gh_time_u', 'cad120', 32, 2, 8, 2, 2.0, {'message_geometry_to_human': True, 'add_time_position': 1, 'time_position_strategy': 'u'}),
    ('grad_mphoi_s2_len_e', 'mphoi', 32, 2, 9, 2, 2.0, {'add_segment_length': 1}),
    ('grad_cad120_nf_len_p', 'cad120', 32, 2, 8, 2, 2.0, {'add_segment_length': 1, 'positional_encoding_style': 'p', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    ('grad_cad120_s2_time_len', 'cad120', 32, 2, 8, 2, 2.0, {'add_segment_length': 1, 'add_time_position': 1}),
    ('grad_mphoi_s2_dist', 'mphoi', 32, 2, 9, 2, 2.0, {'_distances': True}),
    ('grad_cad120_s2_dist', 'cad120', 32, 2, 8, 2, 2.0, {'_distances': True}),
    ('grad_mphoi_s2_gate2', 'mphoi', 32, 2, 9, 2, 2.0, {'discrete_networks_num_layers': 2}),
    ('grad_cad120_nf_gate2_mix', 'cad120', 32, 2, 8, 2, 2.0, {'discrete_networks_num_layers': 2, 'add_time_position': 1, 'time_position_strategy': 'u', 'message_geometry_to_human': True, 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5, 'object_segment_update_strategy': 'coh'}),
    ('grad_mphoi_s2_gate3', 'mphoi', 32, 2, 9, 2, 2.0, {'discrete_networks_num_layers': 3}),
    ('grad_cad120_s2_gate3_sah_u', 'cad120', 32, 2, 8, 2, 2.0, {'discrete_networks_num_layers': 3, 'add_time_position': 1, 'time_position_strategy': 'u', 'object_segment_update_strategy': 'sah'}),
    # no gradient case for discrete_optimization_strategy 'st': the reference's StraightThroughEstimator.backward returns one gradient
    # for two forward inputs and autograd rejects it (distributions.py:39-53) — the reference cannot train with it
    # hidden 512 (the benchmarked width), T = 32: the D=512 BPTT and split-K weight-gradient paths against the reference itself
    ('grad_mphoi_s2_d512', 'mphoi', 512, 8, 32, 2, 1.0),
]


# Under mean pooling every sender weighs 1/n, so a message pre-activation that sits within rounding distance of zero (a ReLU
# knife edge: implementations with a different summation order disagree on its sign) shifts the gradient of its MLP visibly —
# with attention such a sender usually carries a negligible weight.  Found on grad_cad120_s2_mp, seed 300: one
# objects_to_object_message_mlp pre-activation of 5.3e-7 flipped under the GPU's 3xTF32 product and moved the bias gradient by
# 0.7 %.  For these cases the seed search therefore also demands a margin on EVERY ReLU pre-activation of the oracle forward.
RELU_STABLE_CASES = {'grad_mphoi_s2_mp', 'grad_cad120_s2_mp', 'grad_mphoi_s2_dist', 'grad_cad120_s2_dist'}


def _relu_margin(p64, ocfg, batch, hseg, oseg, noise, n_calls, dists=None):
    worst = [float('inf')]
    orig = orc._relu_lin

    def spy(pp, name, x):
        pre = orc._lin(pp, name, x)
        worst[0] = min(worst[0], float(pre.abs().min()))
        return torch.relu(pre)
    orc._relu_lin = spy
    try:
        orc.forward(p64, ocfg, batch['x_human'].double(), batch['x_objects'].double(), batch['objects_mask'].double(),
                    None if hseg is None else hseg.double(), None if oseg is None else oseg.double(),
                    noise.double() if n_calls else None, training=True, steps_per_example=batch['steps_per_example'], distances=dists64(dists))
    finally:
        orc._relu_lin = orig
    return worst[0]


def summarize_grad(g):
    """Full tensor when small, otherwise sums + fixed samples (keeps the fixtures small)."""
    f = g.detach().double().reshape(-1)
    if f.numel() <= 4096:
        return f.numpy()
    idx = torch.linspace(0, f.numel() - 1, 256).long()
    return np.concatenate([[float(f.sum()), float(f.abs().sum()), float((f * f).sum())], f[idx].numpy()])


def run_grad_case(name, shape_name, D, B, T, stage, gain, extra=None):
    shape = pkg.SHAPES[shape_name]
    extra = extra or {}
    kw = pkg.model_kwargs(shape, hidden_size=D, stage=stage, **extra)
    model = select_model('2G-GCN')(**kw)
    pkg.deterministic_fill(model.state_dict(), seed=7, gain=gain)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.train(True)
    misc = dict(impose_segmentation_pattern=1 if stage == 1 else 0,
                segmentation_loss=dict(add=(stage == 2), sigma=4.0 if stage == 2 else 0.0, weight=1.0),
                make_attention_distance_based=bool(extra.get('_distances')))
    feed = select_model_data_feeder('2G-GCN', 'multiple', dataset_name=shape.dataset, **misc)
    criterion, _ = select_loss('2G-GCN', 'multiple', shape.dataset, cfg=Cfg(misc=misc))
    thr = kw['update_segment_threshold']
    for attempt in range(50):
        data_seed, noise_seed = 300 + attempt, 700 + attempt
        batch = pkg.make_batch(shape, B, T, seed=data_seed)
        dists = pkg.make_distances(shape, B, T, seed=data_seed + 5000) if extra.get('_distances') else None
        dist_data = [None, None, None] if dists is None else [None if shape.dataset == 'cad120' else dists[0], dists[1], dists[2]]
        n_calls = orc.num_noise_draws(T, shape.H, shape.O, stage == 1, stage == 1 and shape.dataset == 'cad120',
                                      kw['object_segment_update_strategy'],
                                      kw['discrete_optimization_strategy'] in ('st', 'straight-through'))
        noise = orc.draw_noise(max(n_calls, 1), B, torch.Generator().manual_seed(noise_seed))[:n_calls]
        # margin check on the fp64 oracle (covers object gates that MPHOI does not return)
        p64 = {k: v.double() for k, v in sd.items()}
        ocfg = orc.config_from_kwargs(kw)
        hseg = torch.ones(B, T, shape.H) if stage == 1 else None
        oseg = torch.ones(B, T, shape.O) if (stage == 1 and shape.dataset == 'cad120') else None
        taps = {}
        o64 = orc.forward(p64, ocfg, batch['x_human'].double(), batch['x_objects'].double(), batch['objects_mask'].double(),
                          None if hseg is None else hseg.double(), None if oseg is None else oseg.double(),
                          noise.double() if n_calls else None, training=True, taps=taps,
                          steps_per_example=batch['steps_per_example'], distances=dists64(dists))
        softs = ([o64[1]] if shape.num_classes[1] is None else [o64[2], o64[3]]) if stage == 2 else []
        if oseg is None:
            softs.append(taps['y_oss'])
        margin = 1.0
        for sft in softs:
            margin = min(margin, float((sft - thr).abs().min()))
            if kw['filter_discrete_updates']:
                margin = min(margin, float((sft[:, 1:] - sft[:, :-1]).abs().min()))
        if margin > CASE_MARGIN.get(name, 1e-4) and (name not in RELU_STABLE_CASES or _relu_margin(p64, ocfg, batch, hseg, oseg, noise, n_calls, dists) > 2e-5):
            break
    else:
        raise RuntimeError(f'no safe seed for {name}')
    it = iter(noise)

    def injected(p, temperature=1.0):
        y = torch.log(torch.cat([p, 1.0 - p], -1) + 1e-20) + next(it).to(p)
        return torch.softmax(y / temperature, -1)[:, :1]

    ref_dist.sample_from_gumbel_sigmoid = injected
    model.load_state_dict(sd)
    model.zero_grad()
    data = [batch['x_human'], batch['x_objects'], batch['objects_mask'], None] + dist_data + [batch['steps_per_example']]
    out = feed(model, data)
    tg = pkg.make_targets(shape, batch['lengths'], T, seed=900 + attempt)
    targets = pkg.target_list(shape, tg)
    losses = criterion(out, targets, reduction='mean')
    total = sum(losses)
    total.backward()
    blob = {'meta': np.array([data_seed, noise_seed, 900 + attempt, 7], dtype=np.int64), 'gain': np.array([gain]),
            'loss': np.array([float(total)]), 'losses': np.array([float(l) for l in losses]),
            'weights_checksum': np.array([pkg.state_checksum(sd)])}
    none_keys = []
    for k, prm in model.named_parameters():
        if prm.grad is None:
            none_keys.append(k)
        else:
            blob['grad.' + k] = summarize_grad(prm.grad)
    blob['none_grad_keys'] = np.array(none_keys)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', name + '.npz'), **blob)
    print(f'{name:20s} seeds=({data_seed},{noise_seed}) margin={margin:.2e} loss={float(total):.6f} '
          f'params with grad: {len(blob) - 6}, without: {len(none_keys)}')


if __name__ == '__main__':
    torch.set_num_threads(8)
    only = [a for a in sys.argv[1:] if not a.startswith('--')]       # optional case names: regenerate just those
    if '--grads-only' not in sys.argv:
        for case in CASES:
            if not only or case[0] in only:
                run_case(*case)
    for case in GRAD_CASES:
        if not only or case[0] in only:
            run_grad_case(*case)
