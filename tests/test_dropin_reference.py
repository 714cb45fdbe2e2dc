"""CPU contract test against the UNMODIFIED reference scripts' own plumbing, when /root/reference is present (build container;
skipped on the GPU box): ``install_dropin()`` must make ``vhoi.models.select_model('2G-GCN')`` — what train.py:27 and
predict.py:36 call — return this package's class; the class must construct from the yaml kwargs exactly as train.py:28-34 builds
them; state_dicts must interchange strictly with the reference model in both directions; the reference's own ``gcn_fetcher`` /
``gcn_forward`` (vhoi/data_loading.py:1233-1315) must drive it with the keyword set it sends (the forward itself needs the GPU:
on CPU the call must reach our kernel entry and refuse the CPU tensors, not fail earlier on a signature mismatch)."""
import importlib
import os
import sys
import types

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'vhoi')), reason='reference tree not present on this box')


@pytest.fixture(scope='module')
def ref():
    sys.modules.setdefault('zarr', types.ModuleType('zarr'))      # vhoi/data_loading.py:14 imports zarr for the on-disk loaders only
    if REF not in sys.path:
        sys.path.insert(0, REF)
    models = importlib.import_module('vhoi.models')
    data_loading = importlib.import_module('vhoi.data_loading')
    original = models.TGGCN
    yield types.SimpleNamespace(models=models, data_loading=data_loading, original=original)
    models.TGGCN = original


VARIANTS = [{}, {'add_time_position': 1, 'add_segment_length': 1, 'message_geometry_to_human': True, '_distances': True},
            {'object_segment_update_strategy': 'sah', 'add_time_position': 1, 'time_position_strategy': 'u',
             'positional_encoding_style': 'p', 'discrete_optimization_strategy': 'st', '_distances': True}]


@pytest.mark.parametrize('extra', VARIANTS, ids=['yaml', 'blocks+dist', 'sah+u+p+st+dist'])
@pytest.mark.parametrize('shape_name,stage', [('mphoi', 1), ('mphoi', 2), ('cad120', 2), ('bimanual', 1)])
def test_install_dropin_and_reference_plumbing(shape_name, stage, extra, ref, pkg, synth):
    shape = synth.SHAPES[shape_name]
    if extra.get('object_segment_update_strategy') == 'sah' and shape.H != 1:
        pytest.skip("'sah' needs exactly one human (the reference has no object gate MLP to fall back on)")
    kwargs = synth.model_kwargs(shape, hidden_size=32, stage=stage, **extra)
    ref.models.TGGCN = ref.original
    ref_model = ref.models.select_model('2G-GCN')(**kwargs)
    pkg.install_dropin()
    Model = ref.models.select_model('2G-GCN')                     # train.py:27 / predict.py:36
    assert Model is pkg.TGGCN
    model = Model(**kwargs)                                       # train.py:28-34
    # checkpoints interchange strictly, both ways (train.py:37 / predict.py:43 use strict=False; strict is the stronger statement)
    model.load_state_dict(ref_model.state_dict(), strict=True)
    ref_model.load_state_dict(model.state_dict(), strict=True)
    assert [tuple(v.shape) for v in model.state_dict().values()] == [tuple(v.shape) for v in ref_model.state_dict().values()]
    # the reference's fetcher + feeder with their own kwargs
    batch = synth.make_batch(shape, 2, 6, seed=1)
    tg = synth.target_list(shape, synth.make_targets(shape, batch['lengths'], 6, seed=2))
    zeros = torch.zeros(2, 1)
    hh, ho, oo = synth.make_distances(shape, 2, 6, seed=3) if extra.get('_distances') else (zeros, zeros, zeros)
    dataset = [batch['x_human'], batch['x_objects'], batch['objects_mask'], zeros, zeros if hh is None else hh, ho, oo,
               batch['steps_per_example']] + tg
    misc = dict(impose_segmentation_pattern=1 if stage == 1 else 0, dataset_name=shape.dataset,
                make_attention_distance_based=bool(extra.get('_distances')))
    fetch = ref.data_loading.select_model_data_fetcher('2G-GCN', 'multiple', **misc)
    feed = ref.data_loading.select_model_data_feeder('2G-GCN', 'multiple', **misc)
    data, target = fetch(dataset, device='cpu')
    assert len(target) == len(tg)
    with torch.no_grad():
        want = feed(ref_model, data)                              # the reference model runs on CPU
        with pytest.raises(pkg.abi.TggcnError, match='CUDA device only'):
            feed(model, data)                                     # ours accepts the same call and refuses only the device
    assert len(want) == (6 if shape.num_classes[1] is None else 12)
