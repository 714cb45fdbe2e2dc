"""Rebuild the inputs/weights/noise of a golden case from its recorded seeds (no reference needed)."""
import importlib
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# name -> (shape, D, B, T, stage, train_mode, inspect)   (mirrors oracle/gen_golden.py:CASES)
CASES = {
    'mphoi_s1_eval': ('mphoi', 32, 2, 12, 1, False, False),
    'mphoi_s2_eval': ('mphoi', 32, 3, 14, 2, False, True),
    'mphoi_s2_train_bn': ('mphoi', 32, 2, 11, 2, True, False),
    'mphoi_s2_d64': ('mphoi', 64, 4, 40, 2, False, False),
    'cad120_s1_eval': ('cad120', 32, 2, 10, 1, False, False),
    'cad120_s2_eval': ('cad120', 32, 2, 13, 2, False, False),
    'bimanual_s2_eval': ('bimanual', 16, 2, 9, 2, False, False),
    # model variants beyond the shipped yaml values: trailing dict = constructor overrides
    'mphoi_s2_cat': ('mphoi', 32, 2, 12, 2, False, False, {'cat_level_states': 1}),
    'cad120_s2_cat': ('cad120', 32, 2, 11, 2, False, False, {'cat_level_states': 1}),
    'mphoi_s2_share': ('mphoi', 32, 2, 12, 2, False, False, {'share_level_mlps': 1}),
    'mphoi_s2_mp': ('mphoi', 32, 2, 12, 2, False, False, {'message_aggregation': 'mp'}),
    'cad120_s2_mp': ('cad120', 32, 2, 11, 2, False, False, {'message_aggregation': 'mp'}),
    'mphoi_s2_dot': ('mphoi', 32, 2, 12, 2, False, False, {'attention_style': 'v2'}),
    'cad120_s2_dot': ('cad120', 32, 2, 11, 2, False, False, {'attention_style': 'v2'}),
    'cad120_s2_sah': ('cad120', 32, 2, 11, 2, False, False, {'object_segment_update_strategy': 'sah'}),
    'cad120_nf_sah': ('cad120', 32, 2, 11, 2, False, False, {'object_segment_update_strategy': 'sah', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    'cad120_nf_coh': ('cad120', 32, 2, 11, 2, False, False, {'object_segment_update_strategy': 'coh', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    'mphoi_s2_time_se': ('mphoi', 32, 2, 12, 2, False, False, {'add_time_position': 1, 'time_position_strategy': 's', 'positional_encoding_style': 'e'}),
    'cad120_s2_time_sp': ('cad120', 32, 2, 11, 2, False, False, {'add_time_position': 1, 'time_position_strategy': 's', 'positional_encoding_style': 'p'}),
    'cad120_s2_time_ue': ('cad120', 32, 2, 11, 2, False, False, {'add_time_position': 1, 'time_position_strategy': 'u', 'positional_encoding_style': 'e'}),
    'mphoi_s2_time_up': ('mphoi', 32, 2, 12, 2, False, False, {'add_time_position': 1, 'time_position_strategy': 'u', 'positional_encoding_style': 'p'}),
    'mphoi_s2_st': ('mphoi', 32, 2, 12, 2, False, False, {'discrete_optimization_strategy': 'st'}),
    'cad120_nf_st': ('cad120', 32, 2, 11, 2, False, False, {'discrete_optimization_strategy': 'st', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    'mphoi_s2_gh': ('mphoi', 32, 2, 12, 2, False, False, {'message_geometry_to_human': True}),
    'cad120_s2_gh_time_u': ('cad120', 32, 2, 11, 2, False, False, {'message_geometry_to_human': True, 'add_time_position': 1, 'time_position_strategy': 'u'}),
    'mphoi_s2_len_e': ('mphoi', 32, 2, 12, 2, False, False, {'add_segment_length': 1}),
    'cad120_nf_len_p': ('cad120', 32, 2, 11, 2, False, False, {'add_segment_length': 1, 'positional_encoding_style': 'p', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    'cad120_s2_time_len': ('cad120', 32, 2, 11, 2, False, False, {'add_segment_length': 1, 'add_time_position': 1}),
    'mphoi_s2_dist': ('mphoi', 32, 2, 12, 2, False, False, {'_distances': True}),
    'cad120_s2_dist': ('cad120', 32, 2, 11, 2, False, False, {'_distances': True}),
    'mphoi_s2_gate2': ('mphoi', 32, 2, 12, 2, False, False, {'discrete_networks_num_layers': 2}),
    'cad120_nf_gate2_mix': ('cad120', 32, 2, 11, 2, False, False, {'discrete_networks_num_layers': 2, 'add_time_position': 1, 'time_position_strategy': 'u', 'message_geometry_to_human': True, 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5, 'object_segment_update_strategy': 'coh'}),
    'mphoi_s2_gate3': ('mphoi', 32, 2, 12, 2, False, False, {'discrete_networks_num_layers': 3}),
    'cad120_s2_gate3_sah_u': ('cad120', 32, 2, 11, 2, False, False, {'discrete_networks_num_layers': 3, 'add_time_position': 1, 'time_position_strategy': 'u', 'object_segment_update_strategy': 'sah'}),
}

# BASELINE.json configs[1] itself — what bench.py times (reference outputs; gate margin 2.5e-5).  Kept apart from CASES: the
# tests that loop over CASES demand exact argmax equality and per-kernel taps, this one is compared in tests/test_gpu_fullsize.py.
FULL_CASES = {
    'mphoi_s2_d512_full': ('mphoi', 512, 8, 128, 2, False, False),
}

# name -> (shape, D, B, T, stage[, constructor overrides])   (mirrors oracle/gen_golden.py:GRAD_CASES)
GRAD_CASES = {
    'grad_mphoi_s1': ('mphoi', 32, 2, 9, 1),
    'grad_mphoi_s2': ('mphoi', 32, 3, 10, 2),
    'grad_cad120_s2': ('cad120', 32, 2, 8, 2),
    'grad_mphoi_s2_cat': ('mphoi', 32, 2, 9, 2, {'cat_level_states': 1}),
    'grad_cad120_s2_cat': ('cad120', 32, 2, 8, 2, {'cat_level_states': 1}),
    'grad_mphoi_s2_share': ('mphoi', 32, 2, 9, 2, {'share_level_mlps': 1}),
    # seeds of these two also keep every ReLU pre-activation > 2e-5 from zero (oracle/gen_golden.py: RELU_STABLE_CASES)
    'grad_mphoi_s2_mp': ('mphoi', 32, 2, 9, 2, {'message_aggregation': 'mp'}),
    'grad_cad120_s2_mp': ('cad120', 32, 2, 8, 2, {'message_aggregation': 'mp'}),
    'grad_mphoi_s2_dot': ('mphoi', 32, 2, 9, 2, {'attention_style': 'v2'}),
    'grad_cad120_s2_dot': ('cad120', 32, 2, 8, 2, {'attention_style': 'v2'}),
    'grad_cad120_s2_sah': ('cad120', 32, 2, 8, 2, {'object_segment_update_strategy': 'sah'}),
    'grad_cad120_nf_sah': ('cad120', 32, 2, 8, 2, {'object_segment_update_strategy': 'sah', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    'grad_cad120_nf_coh': ('cad120', 32, 2, 8, 2, {'object_segment_update_strategy': 'coh', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    'grad_mphoi_s2_time_se': ('mphoi', 32, 2, 9, 2, {'add_time_position': 1, 'time_position_strategy': 's', 'positional_encoding_style': 'e'}),
    'grad_cad120_s2_time_sp': ('cad120', 32, 2, 8, 2, {'add_time_position': 1, 'time_position_strategy': 's', 'positional_encoding_style': 'p'}),
    'grad_cad120_s2_time_ue': ('cad120', 32, 2, 8, 2, {'add_time_position': 1, 'time_position_strategy': 'u', 'positional_encoding_style': 'e'}),
    'grad_mphoi_s2_time_up': ('mphoi', 32, 2, 9, 2, {'add_time_position': 1, 'time_position_strategy': 'u', 'positional_encoding_style': 'p'}),
    'grad_mphoi_s2_gh': ('mphoi', 32, 2, 9, 2, {'message_geometry_to_human': True}),
    'grad_cad120_s2_gh_time_u': ('cad120', 32, 2, 8, 2, {'message_geometry_to_human': True, 'add_time_position': 1, 'time_position_strategy': 'u'}),
    'grad_mphoi_s2_len_e': ('mphoi', 32, 2, 9, 2, {'add_segment_length': 1}),
    'grad_cad120_nf_len_p': ('cad120', 32, 2, 8, 2, {'add_segment_length': 1, 'positional_encoding_style': 'p', 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5}),
    'grad_cad120_s2_time_len': ('cad120', 32, 2, 8, 2, {'add_segment_length': 1, 'add_time_position': 1}),
    'grad_mphoi_s2_dist': ('mphoi', 32, 2, 9, 2, {'_distances': True}),
    'grad_cad120_s2_dist': ('cad120', 32, 2, 8, 2, {'_distances': True}),
    'grad_mphoi_s2_gate2': ('mphoi', 32, 2, 9, 2, {'discrete_networks_num_layers': 2}),
    'grad_cad120_nf_gate2_mix': ('cad120', 32, 2, 8, 2, {'discrete_networks_num_layers': 2, 'add_time_position': 1, 'time_position_strategy': 'u', 'message_geometry_to_human': True, 'filter_discrete_updates': 0, 'update_segment_threshold': 0.5, 'object_segment_update_strategy': 'coh'}),
    'grad_mphoi_s2_gate3': ('mphoi', 32, 2, 9, 2, {'discrete_networks_num_layers': 3}),
    'grad_cad120_s2_gate3_sah_u': ('cad120', 32, 2, 8, 2, {'discrete_networks_num_layers': 3, 'add_time_position': 1, 'time_position_strategy': 'u', 'object_segment_update_strategy': 'sah'}),
}

# hidden 512 (the benchmarked width), T = 32.  Kept apart from GRAD_CASES: the reference's own fp32 autograd carries summation
# noise of ~3e-4 of a tensor's largest entry at this size, so tests/test_gpu_fullsize.py compares it with a matching tolerance
# (the tight check at this size is the fp64 oracle, same file).
FULL_GRAD_CASES = {
    'grad_mphoi_s2_d512': ('mphoi', 512, 8, 32, 2),
}


def dist_kwargs(dists, to=None):
    """The three distance kwargs of TGGCN.forward (vhoi/models.py:585) from a (hh, ho, oo) tuple, optionally mapped through `to`."""
    if dists is None:
        return {}
    f = to or (lambda t: t)
    names = ('human_human_distances', 'human_object_distances', 'object_object_distances')
    return {k: (None if d is None else f(d)) for k, d in zip(names, dists)}


def dists64(dists):
    return None if dists is None else tuple(None if d is None else d.double() for d in dists)


def alias_shared_heads(params, extra):
    """share_level_mlps: both names of a shared head must be ONE tensor in a parameter dict used for autograd."""
    if extra.get('share_level_mlps'):
        for k in list(params):
            if '_frame_' in k and k.replace('_frame_', '_') in params:
                params[k] = params[k.replace('_frame_', '_')]
    return params


class GoldenCase:
    def __init__(self, name):
        import tggcn_oracle as orc
        synth = importlib.import_module('2g-gcn_b200.synth')
        self.name = name
        spec = CASES[name] if name in CASES else FULL_CASES[name]
        shape_name, D, B, T, stage, train_mode, inspect = spec[:7]
        self.extra = spec[7] if len(spec) > 7 else {}
        self.shape = synth.SHAPES[shape_name]
        self.D, self.B, self.T, self.stage, self.train_mode, self.inspect = D, B, T, stage, train_mode, inspect
        self.blob = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        data_seed, noise_seed, target_seed, weight_seed = [int(v) for v in self.blob['meta']]
        self.weight_seed, self.gain = weight_seed, float(self.blob['gain'][0])
        self.kwargs = synth.model_kwargs(self.shape, hidden_size=D, stage=stage, **self.extra)
        self.thr = self.kwargs['update_segment_threshold']
        self.batch = synth.make_batch(self.shape, B, T, seed=data_seed)
        # misc.make_attention_distance_based cases: (hh, ho, oo) distances, regenerated from the data seed like the generator did
        self.dists = synth.make_distances(self.shape, B, T, seed=data_seed + 5000) if self.extra.get('_distances') else None
        H, O = self.shape.H, self.shape.O
        self.human_given = stage == 1
        self.objects_given = stage == 1 and self.shape.dataset == 'cad120'
        n_calls = orc.num_noise_draws(T, H, O, self.human_given, self.objects_given,
                                      self.kwargs['object_segment_update_strategy'],
                                      self.kwargs['discrete_optimization_strategy'] in ('st', 'straight-through'))
        self.noise = orc.draw_noise(max(n_calls, 1), B, torch.Generator().manual_seed(noise_seed))[:n_calls]
        self.hseg = torch.ones(B, T, H) if self.human_given else None
        self.oseg = torch.ones(B, T, O) if self.objects_given else None
        self.targets = synth.target_list(self.shape, synth.make_targets(self.shape, self.batch['lengths'], T,
                                                                        seed=target_seed))
        self.outputs = [torch.from_numpy(self.blob[f'out{i}']) for i in range(6 if self.shape.num_classes[1] is None else 12)]
        self.ocfg = orc.config_from_kwargs(self.kwargs)
        # the regenerated inputs must be the bytes the reference saw
        chk = float(self.batch['x_human'].double().sum() + self.batch['x_objects'].double().sum())
        assert abs(chk - float(self.blob['inputs_checksum'][0])) <= 1e-6 * abs(chk), 'synthetic inputs differ from golden run'
        nchk = float(self.noise.double().sum())
        assert abs(nchk - float(self.blob['noise_checksum'][0])) <= 1e-9 * max(1.0, abs(nchk)), 'noise differs from golden run'

    def fill(self, state_dict):
        synth = importlib.import_module('2g-gcn_b200.synth')
        synth.deterministic_fill(state_dict, seed=self.weight_seed, gain=self.gain)
        chk = synth.state_checksum(state_dict)
        ref = float(self.blob['weights_checksum'][0])
        assert abs(chk - ref) <= 1e-9 * max(1.0, abs(ref)), 'regenerated weights differ from golden run'
        return state_dict
